"""The UNMODIFIED reference (ellen660/encodec, installed by __graft_entry__.build() into the git-ignored baseline/_ref with
`pip install --no-index --no-deps --target baseline/_ref /root/reference`) run through its own public API on the host CPU:
`EncodecModel._get_model(...)`, `load_state_dict`, `set_target_bandwidth`, `model(x)` (reference model.py:248-257).

None of this repository's kernels, models or oracle code is on that path; only `encodec_b200.synth` is used, to produce the
same deterministic weights / codebooks / audio that the CUDA arm runs on (so both arms do the same work on the same data).
bench.py's `--impl reference` arm and `cpu_baseline` leg call `throughput()`; tests/test_reference_arm.py checks that the
installed reference reproduces the golden outputs.
"""
from __future__ import annotations

import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "encodec", "model.py"))


def install(source: str = "/root/reference") -> bool:
    """Offline install of the reference into baseline/_ref (build container only: the GPU box has no /root/reference and
    uses the files that travelled with the snapshot). Returns True when baseline/_ref is usable afterwards."""
    import shutil
    import subprocess
    import tempfile
    if available():
        return True
    if not os.path.isdir(source):
        return False
    tmp = tempfile.mkdtemp(prefix="encodec_ref_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(source, src)   # the build writes *.egg-info into the tree; /root/reference is read-only
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", REF_DIR, src]
        subprocess.run(cmd, check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return available()


_MODELS = {}


def build_model(spec, sd, bandwidth, distinct_codebooks: bool = True):
    """The reference model with this repository's synthetic state_dict. Constructor call as SURVEY.md section 8d verified
    it (bins passed explicitly); with distinct codebooks the quantiser's aliased layer list (fork delta D6) is replaced by
    independent VectorQuantization layers -- legal reference behaviour, core_vq.py:397 loops self.layers[:n_q]."""
    import torch
    key = (id(sd), float(bandwidth), distinct_codebooks)
    if key in _MODELS:
        return _MODELS[key]
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    sys.dont_write_bytecode = True
    warnings.filterwarnings("ignore")
    from encodec.model import EncodecModel            # baseline/_ref/encodec
    import quantization.core_vq as core_vq            # registered at top level by the reference's own sys.path handling
    assert os.path.abspath(sys.modules["encodec"].__file__).startswith(os.path.abspath(REF_DIR)), \
        "a different `encodec` package shadows baseline/_ref"
    torch.manual_seed(0)
    m = EncodecModel._get_model(spec.target_bandwidths, spec.sample_rate, spec.channels, causal=spec.causal,
                                model_norm=spec.norm, audio_normalize=spec.normalize, segment=spec.segment, name="unset",
                                ratios=spec.ratios, bins=spec.bins, dimension=spec.dimension).eval()
    if distinct_codebooks:
        n_q = len(m.quantizer.vq.layers)
        m.quantizer.vq.layers = torch.nn.ModuleList([
            core_vq.VectorQuantization(dim=spec.dimension, codebook_size=spec.bins, codebook_dim=spec.dimension, decay=0.99,
                                       kmeans_init=True, kmeans_iters=50, threshold_ema_dead_code=2) for _ in range(n_q)])
    m = m.eval()
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)   # sets inited = 1
    m.set_target_bandwidth(bandwidth)
    _MODELS[key] = m
    return m


def forward(x: np.ndarray, spec, sd, bandwidth, distinct_codebooks: bool = True, threads: int | None = None):
    """(audio, codes) of the reference's forward on CPU."""
    import torch
    torch.set_num_threads(threads or os.cpu_count() or 1)
    m = build_model(spec, sd, bandwidth, distinct_codebooks)
    with torch.no_grad():
        out = m(torch.from_numpy(np.ascontiguousarray(x)))
    return out[0].numpy(), out[1].numpy()


def throughput(spec, sd, bandwidth, clips: int, seconds: float, repeats: int = 1, threads: int | None = None):
    """audio-seconds per second of `model(x)` on `clips` x `seconds` of synthetic audio; returns (value, best seconds)."""
    from encodec_b200 import synth
    x = synth.make_audio(999, clips, spec.channels, int(seconds * spec.sample_rate))
    build_model(spec, sd, bandwidth)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        forward(x, spec, sd, bandwidth, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return clips * seconds / best, best


def rvq_encode(frames: np.ndarray, codebooks: np.ndarray, chunk: int = 65536, threads: int | None = None) -> np.ndarray:
    """BASELINE config 4 on the reference: core_vq.ResidualVectorQuantization.encode on frames [N, D] with distinct codebooks
    [n_q, bins, D]; the reference takes [B, D, T], i.e. frames.T[None], chunked over N for memory."""
    import torch
    torch.set_num_threads(threads or os.cpu_count() or 1)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    sys.dont_write_bytecode = True
    warnings.filterwarnings("ignore")
    from encodec.model import EncodecModel  # noqa: F401  (sets up the reference's top-level module aliases)
    import quantization.core_vq as core_vq
    n_q, bins, dim = codebooks.shape
    rvq = core_vq.ResidualVectorQuantization(num_quantizers=n_q, dim=dim, codebook_size=bins, codebook_dim=dim, decay=0.99,
                                             kmeans_init=True, kmeans_iters=50, threshold_ema_dead_code=2)
    if len({id(l) for l in rvq.layers}) != n_q:   # the fork aliases one layer n_q times (D6): make them independent
        rvq.layers = torch.nn.ModuleList([
            core_vq.VectorQuantization(dim=dim, codebook_size=bins, codebook_dim=dim, decay=0.99, kmeans_init=True,
                                       kmeans_iters=50, threshold_ema_dead_code=2) for _ in range(n_q)])
    rvq = rvq.eval()
    with torch.no_grad():
        for i, layer in enumerate(rvq.layers):
            layer._codebook.embed.copy_(torch.from_numpy(codebooks[i]))
            layer._codebook.inited.fill_(1)
        out = []
        for s in range(0, frames.shape[0], chunk):
            xx = torch.from_numpy(np.ascontiguousarray(frames[s:s + chunk].T))[None]
            out.append(rvq.encode(xx, n_q)[:, 0].numpy())
    return np.concatenate(out, axis=1)


def lm_entropy_round_trip(lm_spec, sd, codes: np.ndarray, threads: int | None = None):
    """The entropy-coded frame loop of the reference (compress.py:66-87 and :125-152) on `codes [K, T]`: its own LMModel one
    step at a time, build_stable_quantized_cdf per codebook, ArithmeticCoder.push / ArithmeticDecoder.pull.
    Returns (seconds to code the frame, seconds to decode it, bytes, decoded codes)."""
    import io
    import torch
    torch.set_num_threads(threads or os.cpu_count() or 1)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    sys.dont_write_bytecode = True
    warnings.filterwarnings("ignore")
    from encodec.model import LMModel
    from encodec.quantization.ac import ArithmeticCoder, ArithmeticDecoder, build_stable_quantized_cdf
    lm = LMModel(lm_spec.n_q, lm_spec.card, dim=lm_spec.dim, num_layers=lm_spec.num_layers, num_heads=lm_spec.num_heads,
                 hidden_scale=lm_spec.hidden_scale, past_context=lm_spec.past_context, max_period=lm_spec.max_period).eval()
    lm.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    frame = torch.from_numpy(np.ascontiguousarray(codes))[None]
    _, K, T = frame.shape
    t0 = time.perf_counter()
    fo = io.BytesIO()
    coder = ArithmeticCoder(fo)
    states, offset = None, 0
    input_ = torch.zeros(1, K, 1, dtype=torch.long)
    for t in range(T):
        with torch.no_grad():
            probas, states, offset = lm(input_, states, offset)
        input_ = 1 + frame[:, :, t: t + 1]
        for k, value in enumerate(frame[0, :, t].tolist()):
            q_cdf = build_stable_quantized_cdf(probas[0, :, k, 0], coder.total_range_bits, check=False)
            coder.push(value, q_cdf)
    coder.flush()
    t1 = time.perf_counter()
    fo.seek(0)
    decoder = ArithmeticDecoder(fo)
    states, offset = None, 0
    input_ = torch.zeros(1, K, 1, dtype=torch.long)
    out = torch.zeros(1, K, T, dtype=torch.long)
    for t in range(T):
        with torch.no_grad():
            probas, states, offset = lm(input_, states, offset)
        code_list = []
        for k in range(K):
            q_cdf = build_stable_quantized_cdf(probas[0, :, k, 0], decoder.total_range_bits, check=False)
            code = decoder.pull(q_cdf)
            if code is None:
                raise EOFError("The stream ended sooner than expected.")
            code_list.append(code)
        out[0, :, t] = torch.tensor(code_list, dtype=torch.long)
        input_ = 1 + out[:, :, t: t + 1]
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, len(fo.getvalue()), out[0].numpy()
