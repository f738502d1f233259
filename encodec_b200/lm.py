"""Language model + arithmetic coder of the entropy-coded ``.ecdc`` stream.

``LMModel`` mirrors the reference class (model.py:45-83): same constructor, same ``state_dict`` keys (the transformer layers
are plain ``nn.TransformerEncoderLayer`` parameter holders, like the reference's subclass transformer.py:30), same
``forward(indices, states, offset) -> (probas [B, card, K, T], states, offset + T)`` streaming contract. The arithmetic
runs in csrc/lm.cu; there is no CPU path. On top of the reference's step-by-step API:

* ``coder_ranges(codes)``  -- compression: every step of every frame in ONE batched pass, returning for each symbol the two
  quantised-cdf values ``ArithmeticCoder.push`` reads (ac.py:143-144);
* ``decode_frame(data, first_byte, K, T)`` -- decompression: the whole step loop (LM step -> cdfs -> ``ArithmeticDecoder.pull``
  -> next input) enqueued on the device, one host synchronisation per frame instead of K ``.item()`` calls per step;
* ``ac_encode`` / ``ac_decode`` -- the coder on the host (C++, ac_core.h), bit-exact with the reference's.
"""
from __future__ import annotations

import ctypes as C
import typing as tp

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _native as nat

TOTAL_RANGE_BITS = 24      # ArithmeticCoder's default (ac.py:99), what compress.py uses

_AC_ERRORS = {1: "The stream ended sooner than expected.", 2: "Binary search failed",
              3: "arithmetic decoder: range representation exceeds 62 bits", 4: "arithmetic decoder: invalid cdf"}


# ---------------------------------------------------------------------------------------------------------------------
# host coder (no GPU involved)
# ---------------------------------------------------------------------------------------------------------------------
def ac_encode(sym_ranges: np.ndarray, total_range_bits: int = TOTAL_RANGE_BITS) -> bytes:
    """ArithmeticCoder.push for every row of ``sym_ranges`` ([n, 2] = (cdf[s - 1] or 0, cdf[s])) + flush (ac.py:127-166)."""
    r = np.ascontiguousarray(sym_ranges, dtype=np.int32).reshape(-1, 2)
    out = np.empty(4 * r.shape[0] + 16, dtype=np.uint8)
    n_out = C.c_int64(0)
    nat.check(nat.lib.ecb_ac_encode(r.ctypes.data, r.shape[0], total_range_bits, out.ctypes.data, out.size, C.byref(n_out)))
    return out[: n_out.value].tobytes()


def ac_decode(data: bytes, cdfs: np.ndarray, total_range_bits: int = TOTAL_RANGE_BITS) -> tp.Tuple[np.ndarray, int]:
    """ArithmeticDecoder.pull against each row of ``cdfs`` ([n, card] int) -> (symbols [n], bytes consumed); raises
    EOFError / RuntimeError with the reference's messages."""
    c = np.ascontiguousarray(cdfs, dtype=np.int32)
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    syms = np.empty(c.shape[0], dtype=np.int32)
    used = C.c_int64(0)
    rc = nat.lib.ecb_ac_decode(buf.ctypes.data if buf.size else None, buf.size, c.ctypes.data, c.shape[0], c.shape[1],
                               total_range_bits, syms.ctypes.data, C.byref(used))
    if rc != 0:
        msg = nat.last_error()
        raise (EOFError if "ended sooner" in msg else RuntimeError)(msg)
    return syms.astype(np.int64), int(used.value)


def ac_decode_device(data: torch.Tensor, cdfs: torch.Tensor, total_range_bits: int = TOTAL_RANGE_BITS):
    """The device decoder of ``LMModel.decode_frame`` alone: uint8 stream + int32 ``cdfs [n, card]`` (CUDA) ->
    ``(symbols [n] int64, bytes consumed)``."""
    nat.require_cuda(data, "data", torch.uint8)
    nat.require_cuda(cdfs, "cdfs", torch.int32)
    c = cdfs.contiguous()
    syms = torch.zeros(c.shape[0], dtype=torch.int64, device=c.device)
    result = torch.zeros(8, dtype=torch.int64, device=c.device)
    with torch.cuda.device(c.device):
        nat.check(nat.lib.ecb_ac_decode_device(nat.ptr(data), data.numel(), nat.ptr(c), c.shape[0], c.shape[1], total_range_bits,
                                               nat.ptr(syms), nat.ptr(result), nat.stream_ptr(c.device)))
    status, used = (int(v) for v in result[:2].cpu().tolist())
    if status != 0:
        raise (EOFError if status == 1 else RuntimeError)(_AC_ERRORS.get(status, f"arithmetic decoder status {status}"))
    return syms, used


def quantized_cdf(pdf: torch.Tensor, total_range_bits: int = TOTAL_RANGE_BITS) -> torch.Tensor:
    """build_stable_quantized_cdf(pdf, total_range_bits, check=False) (ac.py:18-53) for every row of a CUDA float32
    ``pdf [..., card]`` -> int32 cdf of the same shape, bit-exact with the reference's CPU float32 arithmetic."""
    nat.require_cuda(pdf, "pdf")
    p = pdf.contiguous()
    out = torch.empty(p.shape, dtype=torch.int32, device=p.device)
    with torch.cuda.device(p.device):
        nat.check(nat.lib.ecb_quantized_cdf(nat.ptr(p), p.numel() // p.shape[-1], p.shape[-1], total_range_bits, nat.ptr(out),
                                            nat.stream_ptr(p.device)))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# model
# ---------------------------------------------------------------------------------------------------------------------
class StreamingTransformerEncoder(nn.Module):
    """Parameter holder with the reference's layout (modules/transformer.py:62-97): ``norm_in`` + ``layers``."""

    def __init__(self, dim, hidden_scale: float = 4., num_heads: int = 8, num_layers: int = 5, max_period: float = 10000,
                 past_context: int = 1000, gelu: bool = True, norm_in: bool = True, dropout: float = 0., **kwargs):
        super().__init__()
        assert dim % num_heads == 0
        if not gelu or not norm_in or dropout != 0. or kwargs:
            raise NotImplementedError("encodec_b200: the LM kernels implement gelu=True, norm_in=True, dropout=0 and the "
                                      "default (post-norm) nn.TransformerEncoderLayer only")
        self.dim, self.num_heads, self.num_layers = dim, num_heads, num_layers
        self.hidden_dim = int(dim * hidden_scale)
        self.max_period, self.past_context = max_period, past_context
        self.norm_in = nn.LayerNorm(dim)
        self.layers = nn.ModuleList([
            nn.TransformerEncoderLayer(dim, num_heads, self.hidden_dim, activation=F.gelu, batch_first=True, dropout=0.)
            for _ in range(num_layers)])


class LMState:
    """The reference's ``states`` list (one input history per layer, transformer.py:99-118), kept as the projected keys /
    values of every step so far: float32 [layers, B, capacity + 1, 2 dim]."""

    def __init__(self, cache: torch.Tensor, capacity: int, steps: int):
        self.cache, self.capacity, self.steps = cache, capacity, steps


class LMModel(nn.Module):
    """Language model estimating the probabilities of each codebook entry (reference model.py:45-83)."""

    def __init__(self, n_q: int = 32, card: int = 1024, dim: int = 200, **kwargs):
        super().__init__()
        self.card, self.n_q, self.dim = card, n_q, dim
        self.transformer = StreamingTransformerEncoder(dim=dim, **kwargs)
        self.emb = nn.ModuleList([nn.Embedding(card + 1, dim) for _ in range(n_q)])
        self.linears = nn.ModuleList([nn.Linear(dim, card) for _ in range(n_q)])
        self.__dict__["_h"] = None
        self.__dict__["_sig"] = None

    # ---- native handle -------------------------------------------------------------------------------------------
    def _signature(self):
        return tuple((t.data_ptr(), t._version, t.device) for t in self.parameters())

    def __del__(self):
        h = self.__dict__.get("_h")
        if h is not None and h.value:
            try:
                nat.lib.ecb_lm_destroy(h)
            except Exception:
                pass

    def native(self) -> C.c_void_p:
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError("encodec_b200: LMModel parameters must live on a CUDA device (no CPU fallback); call .cuda() first")
        sig = self._signature()
        if self._h is None or sig != self._sig:
            if self._h is not None:
                nat.lib.ecb_lm_destroy(self._h)
                self.__dict__["_h"] = None
            tr = self.transformer
            spec = nat.EcbLmSpec(self.n_q, self.card, self.dim, tr.num_layers, tr.num_heads, tr.hidden_dim, tr.past_context,
                                 float(tr.max_period))
            h = C.c_void_p()
            nat.check(nat.lib.ecb_lm_create(C.byref(spec), C.byref(h)))
            with torch.cuda.device(p.device):
                st = nat.stream_ptr(p.device)
                keep = []
                for key, t in self.state_dict().items():
                    t = t.detach().to(device=p.device, dtype=torch.float32).contiguous()
                    keep.append(t)
                    nat.check(nat.lib.ecb_lm_load_tensor(h, key.encode(), nat.ptr(t), t.numel(), st))
                # the divisors of create_sin_embedding exactly as the reference's framework rounds them (transformer.py:22-23)
                half = self.dim // 2
                adim = torch.arange(half).view(1, 1, -1)
                div = (tr.max_period ** (adim / (half - 1))).to(torch.float32).reshape(-1).contiguous().numpy()
                nat.check(nat.lib.ecb_lm_finalize(h, div.ctypes.data, st))
                torch.cuda.current_stream(p.device).synchronize()
            self.__dict__["_h"] = h
            self.__dict__["_sig"] = sig
        return self._h

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def _new_cache(self, n_items: int, capacity: int) -> torch.Tensor:
        nbytes = int(nat.lib.ecb_lm_cache_bytes(self.native(), n_items, capacity))
        return torch.empty(nbytes // 4, dtype=torch.float32, device=self.device)

    def _run(self, tokens: torch.Tensor, are_codes: bool, t0: int, n_t: int, cache: torch.Tensor, capacity: int,
             probas=None, cdf=None, sym_ranges=None) -> None:
        """tokens [B, K, *] int64 CUDA."""
        nat.require_cuda(tokens, "indices", torch.int64)
        b, k = tokens.shape[0], tokens.shape[1]
        h = self.native()
        with torch.cuda.device(tokens.device):
            ws_bytes = int(nat.lib.ecb_lm_workspace_bytes(h, b * n_t, k))
            ws = nat.shared_workspace(tokens.device, ws_bytes)
            nat.check(nat.lib.ecb_lm_forward(h, nat.ptr(tokens), tokens.stride(0), tokens.stride(1), tokens.stride(2),
                                             1 if are_codes else 0, b, k, t0, n_t, nat.ptr(cache), capacity, nat.ptr(probas),
                                             nat.ptr(cdf), nat.ptr(sym_ranges), nat.ptr(ws), ws.numel(),
                                             nat.stream_ptr(tokens.device)))

    # ---- the reference's streaming API -----------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, indices: torch.Tensor, states: tp.Optional[LMState] = None, offset: int = 0):
        """``indices [B, K, T]``: 1 + the codes of the previous step, 0 where there is none (model.py:67-72). Returns
        ``(probabilities [B, card, K, T], new_states, offset + T)``; pass ``new_states`` / ``offset`` back for the next steps."""
        b, k, t = indices.shape
        if states is None:
            if offset != 0:
                raise ValueError("encodec_b200: a fresh LM state starts at offset 0")
            cap = max(256, t)
            states = LMState(self._new_cache(b, cap), cap, 0)
        if offset != states.steps:
            raise ValueError(f"encodec_b200: offset {offset} does not continue the state ({states.steps} steps so far)")
        if states.steps + t > states.capacity:
            cap = max(2 * states.capacity, states.steps + t)
            new = self._new_cache(b, cap)
            layers, d2 = self.transformer.num_layers, 2 * self.dim
            new.view(layers, b, cap + 1, d2)[:, :, : states.capacity + 1] = states.cache.view(layers, b, states.capacity + 1, d2)
            states = LMState(new, cap, states.steps)
        probas = torch.empty((b, t, k, self.card), dtype=torch.float32, device=indices.device)
        self._run(indices, False, states.steps, t, states.cache, states.capacity, probas=probas)
        return probas.permute(0, 3, 2, 1), LMState(states.cache, states.capacity, states.steps + t), offset + t

    # ---- whole frames ----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def frame_outputs(self, codes: torch.Tensor, probas: bool = False, cdf: bool = False, sym_ranges: bool = True):
        """codes [n_frames, K, T] int64 (CUDA): every frame is an independent stream starting from an empty state
        (compress.py:66-69). Returns a dict with the requested ``probas`` float32 / ``cdf`` int32 [n_frames, T, K, card] and
        ``sym_ranges`` int32 [n_frames, T, K, 2]."""
        n, k, t = codes.shape
        dev = codes.device
        cache = self._new_cache(n, t)
        out = {}
        if probas:
            out["probas"] = torch.empty((n, t, k, self.card), dtype=torch.float32, device=dev)
        if cdf:
            out["cdf"] = torch.empty((n, t, k, self.card), dtype=torch.int32, device=dev)
        if sym_ranges:
            out["sym_ranges"] = torch.empty((n, t, k, 2), dtype=torch.int32, device=dev)
        self._run(codes, True, 0, t, cache, t, probas=out.get("probas"), cdf=out.get("cdf"), sym_ranges=out.get("sym_ranges"))
        return out

    def coder_ranges(self, codes: torch.Tensor) -> torch.Tensor:
        return self.frame_outputs(codes)["sym_ranges"]

    @torch.no_grad()
    def encode_frames(self, codes: torch.Tensor) -> tp.List[bytes]:
        """The bytes ``ArithmeticCoder`` writes for each frame of ``codes [n_frames, K, T]`` (compress.py:66-87)."""
        r = self.coder_ranges(codes).cpu().numpy()
        return [ac_encode(r[i]) for i in range(r.shape[0])]

    @torch.no_grad()
    def decode_frame(self, data: torch.Tensor, first_byte: int, n_codebooks: int, n_steps: int) -> tp.Tuple[torch.Tensor, int]:
        """data: uint8 CUDA tensor holding the stream; the frame's coder starts at ``first_byte``. Returns
        ``(codes [K, T] int64, index of the first byte after this frame)``."""
        nat.require_cuda(data, "data", torch.uint8)
        dev = data.device
        h = self.native()
        codes = torch.zeros((n_codebooks, n_steps), dtype=torch.int64, device=dev)
        result = torch.zeros(2, dtype=torch.int64, device=dev)
        cache = self._new_cache(1, n_steps)
        with torch.cuda.device(dev):
            ws_bytes = int(nat.lib.ecb_lm_workspace_bytes(h, 1, n_codebooks))
            ws = nat.shared_workspace(dev, ws_bytes)
            nat.check(nat.lib.ecb_lm_decode_frame(h, nat.ptr(data), data.numel(), first_byte, n_codebooks, n_steps,
                                                  nat.ptr(codes), nat.ptr(cache), n_steps, nat.ptr(result), nat.ptr(ws),
                                                  ws.numel(), nat.stream_ptr(dev)))
        status, end = (int(v) for v in result.cpu().tolist())
        if status != 0:
            raise (EOFError if status == 1 else RuntimeError)(_AC_ERRORS.get(status, f"arithmetic decoder status {status}"))
        return codes, end
