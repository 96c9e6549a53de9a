"""Encoder output epilogue on the GPU: masked mean pooling (+ L2 normalise) in one kernel.

Replaces the tail of `SentenceTransformer.encode` (reference call sites src/retrieval.py:98,
src/create_embeddings.py:97-101): sentence-transformers `Pooling(pooling_mode_mean_tokens)`
followed, for e5, by `Normalize`.  Tensor hand-off only: the transformer itself is unchanged and
its last_hidden_state never leaves the device; the pooled embeddings can go straight into
`FlatIndex.search` / `.add` without a host hop.
"""
from __future__ import annotations

import ctypes

from . import _lib
from ._lib import check
from .flat import _torch_dtype_code


def mean_pool_normalize(hidden, attention_mask, normalize: bool = False):
    """hidden: CUDA tensor [B, T, H] (fp32/fp16/bf16); attention_mask: CUDA tensor [B, T] (any
    integer/bool dtype).  Returns a CUDA float32 tensor [B, H].  Asynchronous on the current stream."""
    import torch
    if not (hasattr(hidden, "is_cuda") and hidden.is_cuda):
        raise _lib.PrsError(_lib.ECUDA, "mean_pool_normalize needs CUDA tensors: there is no CPU fallback")
    if hidden.dim() != 3 or attention_mask.dim() != 2 or tuple(attention_mask.shape) != tuple(hidden.shape[:2]):
        raise _lib.PrsError(_lib.EINVAL, f"shape mismatch: hidden {tuple(hidden.shape)} mask {tuple(attention_mask.shape)}")
    hidden = hidden.contiguous()
    mask = attention_mask.to(device=hidden.device, dtype=torch.int64).contiguous()
    B, T, H = (int(s) for s in hidden.shape)
    out = torch.empty((B, H), dtype=torch.float32, device=hidden.device)
    st = torch.cuda.current_stream(hidden.device).cuda_stream
    check(_lib.lib().prs_pool_norm(ctypes.c_void_p(hidden.data_ptr()), _torch_dtype_code(hidden),
                                   ctypes.c_void_p(mask.data_ptr()), B, T, H, 1 if normalize else 0,
                                   ctypes.c_void_p(out.data_ptr()), int(hidden.device.index or 0), ctypes.c_void_p(st)))
    return out


class FusedPoolingEncoder:
    """Minimal `SentenceTransformer.encode`-shaped wrapper around an unchanged HF transformer:
    tokenizer -> model forward (stock PyTorch) -> fused pool/normalise kernel.  `dense` is an
    optional torch module applied after pooling (distiluse: Linear 768->512 + tanh)."""

    def __init__(self, model, tokenizer, normalize: bool = False, dense=None, max_seq_length: int = 128, device="cuda"):
        self.model, self.tokenizer, self.normalize, self.dense = model, tokenizer, normalize, dense
        self.max_seq_length, self.device = max_seq_length, device

    def encode_device(self, sentences):
        import torch
        enc = self.tokenizer(list(sentences), padding=True, truncation=True, max_length=self.max_seq_length, return_tensors="pt")
        enc = {k: v.to(self.device) for k, v in enc.items()}
        with torch.no_grad():
            hidden = self.model(**enc).last_hidden_state
            if self.dense is None:
                return mean_pool_normalize(hidden, enc["attention_mask"], self.normalize)
            pooled = mean_pool_normalize(hidden, enc["attention_mask"], False)
            out = self.dense(pooled)
            return torch.nn.functional.normalize(out, p=2, dim=1) if self.normalize else out

    def encode(self, sentences, device=None, **_kw):
        return self.encode_device(sentences).float().cpu().numpy()
