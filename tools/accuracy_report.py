import os, sys, json
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from tests import golden_cases as gc
from tests import util_gpu as ug
from oracle import encodec_oracle as orc
for name in gc.MODEL_CASES:
    case = gc.load_model_case(name)
    spec = case["spec"]
    m = ug.build_model(spec, case["sd"], case["bandwidth"], case["distinct"])
    x = torch.from_numpy(case["x"]).cuda()
    with torch.no_grad():
        audio, codes, _, _ = m(x)
    a = audio.cpu().numpy(); ref = case["audio"]
    d = a - ref
    print(name, "split", os.environ.get("ECB_DEC_SPLIT","3"), "audio max-abs %.3e rms %.3e ref-rms %.3e" % (np.abs(d).max(), np.sqrt((d**2).mean()), np.sqrt((ref**2).mean())),
          "codes mismatch", int((codes.cpu().numpy() != case["codes"]).sum()), "of", case["codes"].size)
    n_q = case["codes"].shape[1]
    cbs = orc.codebooks_from_state_dict(case["sd"], n_q)
    sc = orc.score_codes(gc.frames_of(case["emb"]), cbs, np.transpose(case["codes"], (1, 0, 2)).reshape(n_q, -1),
                         np.transpose(codes.cpu().numpy(), (1, 0, 2)).reshape(n_q, -1))
    emb = m.encoder(x if not spec.normalize else x[:, :, :spec.segment_length]).cpu().numpy() if not spec.normalize else None
    if emb is not None:
        e = emb - case["emb"]
        print("   emb max-abs err %.3e (rel to max %.3e)" % (np.abs(e).max(), np.abs(e).max() / np.abs(case["emb"]).max()))
    print("   teacher-forced score", sc)
