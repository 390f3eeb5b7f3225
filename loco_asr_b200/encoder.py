"""Host-side mirror of the reference's encoder call surface, backed by the C-ABI CUDA library.

Reference call sites this class is a drop-in for (speech_text/extract_speecht5_base_embeddings_slurp.py):

    model = SpeechT5ForSpeechToText.from_pretrained("microsoft/speecht5_asr").to(device)      # :98
    model.speecht5.encoder.wrapped_encoder.load_state_dict(encoder_state_dict)                # :99
    model.speecht5.encoder.prenet.load_state_dict(speech_prenet_state_dict)                   # :100
    out = model.speecht5.encoder(**audios)            # audios = {input_values, attention_mask}   :108
    embeddings = out.last_hidden_state.cpu().detach().numpy()                                 # :109

``LocoSpeechT5Encoder`` plays the role of ``model.speecht5.encoder`` (HF ``SpeechT5EncoderWithSpeechPrenet``,
modeling_speecht5.py:1341-1374): same keyword arguments, same ``.last_hidden_state`` ([B, T_max, 768] fp32),
same ``.prenet`` / ``.wrapped_encoder`` ``load_state_dict`` entry points and checkpoint key names
(map_speecht5_hf.py:34-168).  Extra: ``.pooled`` (mean over each utterance's own frames) and the var-len
fast paths ``encode_packed`` / ``encode_host`` that never build a padded tensor.

Semantics note: every utterance is encoded as if alone (no padding leaks into GroupNorm or
the positional conv), which is what the reference computes for equal-length batches bit-exactly.

PyTorch is used here only for device memory, streams and tensor plumbing; all arithmetic runs in
``libloco_asr.so``.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Mapping, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .config import LocoSpeechT5Config

_TORCH2LOCO = {torch.float32: _lib.LOCO_F32, torch.float16: _lib.LOCO_F16, torch.bfloat16: _lib.LOCO_BF16,
               torch.float64: _lib.LOCO_F64}


@dataclass
class LocoEncoderOutput:
    """Mirrors HF ``BaseModelOutput`` for the fields the reference reads, plus the pooled embedding."""
    last_hidden_state: Optional[torch.Tensor]   # f32[B, T_max, 768], rows past an utterance's length are 0
    pooled: torch.Tensor                        # f32[B, 768]
    frame_lengths: torch.Tensor                 # i64[B] (cpu)
    hidden_states = None
    attentions = None

    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]


class _SubmoduleShim:
    """``encoder.prenet`` / ``encoder.wrapped_encoder``: accepts the sub-module state dicts the reference
    pickles with map_speecht5_hf.py (keys stripped of their 3 leading components, :94-99 / :157-168)."""

    def __init__(self, owner: "LocoSpeechT5Encoder", prefix: str):
        self._owner, self._prefix = owner, prefix

    def load_state_dict(self, state_dict: Mapping[str, torch.Tensor], strict: bool = True):
        self._owner._ingest({self._prefix + k: v for k, v in state_dict.items()})
        return "<All keys matched successfully>"


class LocoPlan:
    """A batch geometry made ready to launch (``loco_plan_create``): encoding with it is a pure enqueue (kernels and memset
    nodes only), so the call can be captured into a CUDA graph and replayed after refilling the same input buffer."""

    def __init__(self, owner: "LocoSpeechT5Encoder", handle, kind: int, lengths: np.ndarray):
        self._owner, self._p, self.kind, self.lengths = owner, handle, kind, lengths
        n = len(lengths)
        self.frames = np.zeros(n, dtype=np.int32)
        self.rows = np.zeros(n, dtype=np.int32)
        total, ws = C.c_int64(), C.c_size_t()
        rc = owner._lib.loco_plan_info(handle, self.frames.ctypes.data, self.rows.ctypes.data, C.byref(total), C.byref(ws))
        _lib.check(owner._lib, owner._h, rc, "loco_plan_info")
        self.total_frames, self.workspace_bytes = int(total.value), int(ws.value)

    def close(self):
        if self._p is not None and getattr(self._owner, "_h", None):
            self._owner._lib.loco_plan_destroy(self._owner._h, self._p)
        self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LocoSpeechT5Encoder:
    def __init__(self, config=None, device: "torch.device | str | int" = "cuda:0", debug: bool = False):
        self.config = LocoSpeechT5Config.from_hf(config) if config is not None else LocoSpeechT5Config()
        self.config.validate()
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise _lib.LocoError(f"LocoSpeechT5Encoder runs on CUDA (sm_100a) only, got device {dev}; there is no CPU fallback")
        if not torch.cuda.is_available():
            raise _lib.LocoError("no CUDA device is visible; the B200 extension cannot run and there is no CPU fallback")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self._lib = _lib.load(debug=debug)       # debug: the LOCO_DEBUG build with the cross-check kernels (tests, tools/)
        cc = _lib.LocoConfigC()
        self._lib.loco_default_config(C.byref(cc))
        cc.encoder_layers = self.config.encoder_layers
        cc.max_speech_positions = self.config.max_speech_positions
        cc.pad_token_id = self.config.pad_token_id
        cc.layer_norm_eps = self.config.layer_norm_eps
        h = C.c_void_p()
        rc = self._lib.loco_create(C.byref(cc), self.device.index, C.byref(h))
        _lib.check(self._lib, None, rc, "loco_create")
        self._h = h
        self._finalized = False
        self._text_vocab = 0
        self._workspace: Optional[torch.Tensor] = None
        self.prenet = _SubmoduleShim(self, "prenet.")
        self.wrapped_encoder = _SubmoduleShim(self, "wrapped_encoder.")

    # ------------------------------------------------------------------ lifecycle / HF-module look-alikes
    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.loco_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def eval(self):
        return self

    def to(self, device=None, *a, **k):
        if device is not None and torch.device(device).type != "cuda":
            raise _lib.LocoError("LocoSpeechT5Encoder cannot move off the GPU (no CPU fallback)")
        return self

    def load_state_dict(self, state_dict: Mapping[str, torch.Tensor], strict: bool = True):
        """Full-encoder dict: HF keys with or without the ``speecht5.encoder.`` prefix."""
        self._ingest(dict(state_dict))
        return "<All keys matched successfully>"

    @classmethod
    def from_state_dict(cls, state_dict, config=None, device="cuda:0", debug: bool = False) -> "LocoSpeechT5Encoder":
        enc = cls(config, device, debug=debug)
        enc.load_state_dict(state_dict)
        enc.finalize()
        return enc

    def _ingest(self, sd: Dict[str, torch.Tensor]):
        if self._finalized:
            raise _lib.LocoError("weights are already finalized; create a new encoder to load another checkpoint")
        for k, v in sd.items():
            if k.endswith("prenet.embed_tokens.weight"):
                self._text_vocab = int(v.shape[0])
            if not (k.startswith("prenet.") or k.startswith("wrapped_encoder.") or k.startswith("speecht5.encoder.")
                    or k.startswith("encoder.")):
                continue  # decoder / head tensors of a full-model checkpoint are not on this path
            t = v.detach()
            if t.dtype not in _TORCH2LOCO:
                t = t.float()
            t = t.cpu().contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            if t.dtype == torch.bfloat16:
                ptr = t.view(torch.int16).numpy().ctypes.data
            else:
                ptr = t.numpy().ctypes.data
            rc = self._lib.loco_load_tensor(self._h, k.encode(), C.c_void_p(ptr), shape, t.dim(), _TORCH2LOCO[t.dtype])
            _lib.check(self._lib, self._h, rc, f"loco_load_tensor({k})")

    def finalize(self):
        if not self._finalized:
            rc = self._lib.loco_finalize_weights(self._h)
            _lib.check(self._lib, self._h, rc, "loco_finalize_weights")
            self._finalized = True
        return self

    # ------------------------------------------------------------------ geometry
    def plan(self, n_samples: Sequence[int]):
        ns = np.ascontiguousarray(np.asarray(n_samples, dtype=np.int32))
        n = int(ns.shape[0])
        frames = np.zeros(n, dtype=np.int32)
        rows = np.zeros(n, dtype=np.int32)
        total = C.c_int64()
        ws = C.c_size_t()
        rc = self._lib.loco_plan(self._h, ns.ctypes.data, n, frames.ctypes.data, rows.ctypes.data, C.byref(total), C.byref(ws))
        _lib.check(self._lib, self._h, rc, "loco_plan")
        return {"frames": frames, "rows": rows, "total_frames": int(total.value), "workspace_bytes": int(ws.value)}

    def _get_workspace(self, nbytes: int) -> torch.Tensor:
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = None
            self._workspace = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device=self.device)
        return self._workspace       # any alignment: the library rounds the base up itself

    # ------------------------------------------------------------------ plans: pure-enqueue encodes (CUDA-graph capturable)
    def make_plan(self, lengths: Sequence[int], text: bool = False) -> LocoPlan:
        """``loco_plan_create`` for one batch geometry (samples per utterance, or tokens per text).  Synchronous; do it
        outside hot loops and outside stream captures."""
        self.finalize()
        ls = np.ascontiguousarray(np.asarray(lengths, dtype=np.int32))
        p = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self._lib.loco_plan_create(self._h, 1 if text else 0, ls.ctypes.data, int(ls.shape[0]), C.byref(p))
        _lib.check(self._lib, self._h, rc, "loco_plan_create")
        return LocoPlan(self, p, 1 if text else 0, ls)

    def encode_planned(self, plan: LocoPlan, inp: torch.Tensor, pooled: torch.Tensor, workspace: torch.Tensor,
                       hidden: Optional[torch.Tensor] = None):
        """``loco_encode_planned``: every buffer is the caller's (``inp`` f32 waveforms / i32 tokens [sum lengths], ``pooled``
        f32 [B, 768], ``workspace`` uint8 [>= plan.workspace_bytes], optional ``hidden`` f32 [plan.total_frames, 768]).
        Enqueues on the current stream and returns; nothing is allocated, copied from the host or synchronised."""
        if inp.numel() != int(plan.lengths.sum()) or inp.device != self.device:
            raise _lib.LocoError("encode_planned: the input does not match the plan")
        with torch.cuda.device(self.device):
            rc = self._lib.loco_encode_planned(self._h, plan._p, inp.data_ptr(), pooled.data_ptr(),
                                               hidden.data_ptr() if hidden is not None else None, workspace.data_ptr(),
                                               workspace.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(self._lib, self._h, rc, "loco_encode_planned")
        return pooled

    def sync_check(self):
        """``loco_sync_check``: synchronise the current stream and raise on any asynchronous CUDA error."""
        with torch.cuda.device(self.device):
            rc = self._lib.loco_sync_check(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(self._lib, self._h, rc, "loco_sync_check")

    # ------------------------------------------------------------------ var-len fast path (device buffers)
    # ------------------------------------------------------------------ classifier head as the encoder's epilogue
    _POOL_METHODS = {"average": 0, "max": 1, "attention": 2, "self_attention": 2}

    def set_head(self, head=None, method: Optional[str] = None, q=None, weight=None, bias=None):
        """Fuse the intent classifier (``speech_text/intent_classifier.py:14-50``) into the last encoder kernel.
        ``head`` is a ``loco_asr_b200.head.IntentHead`` (or pass method / q [1,768] / weight [C,768] / bias [C]).  Afterwards
        ``encode_packed(..., with_head=True)`` / ``encode_text_packed(..., with_head=True)`` also return the method's pooled
        vector and the logits."""
        if head is not None:
            method, q, weight, bias = head.method, head.q, head.weight, head.bias
        if method not in self._POOL_METHODS:
            raise ValueError(f"unknown pooling method {method!r} (reference: average / max / self_attention)")

        def host(t, n):
            if t is None:
                return None
            a = np.ascontiguousarray(torch.as_tensor(t).detach().to("cpu", torch.float32).reshape(-1).numpy())
            if a.size != n:
                raise _lib.LocoError(f"head tensor has {a.size} elements, expected {n}")
            return a

        hid = self.config.hidden_size
        n_classes = int(weight.shape[0]) if weight is not None else 0
        qh, wh, bh = host(q, hid), host(weight, n_classes * hid), host(bias, n_classes)
        ptr = lambda a: a.ctypes.data if a is not None else None
        with torch.cuda.device(self.device):
            rc = self._lib.loco_set_head(self._h, self._POOL_METHODS[method], ptr(qh), ptr(wh), ptr(bh), n_classes)
        _lib.check(self._lib, self._h, rc, "loco_set_head")
        self._head_classes = n_classes
        self._head_method = method

    def _head_outputs(self, n: int, with_head: bool):
        """Allocate the head's outputs for one encode call and point the library at them."""
        if not with_head:
            return None, None
        if getattr(self, "_head_method", None) is None:
            raise _lib.LocoError("with_head=True before set_head()")
        hp = torch.empty(n, self.config.hidden_size, dtype=torch.float32, device=self.device)
        lg = torch.empty(n, self._head_classes, dtype=torch.float32, device=self.device) if self._head_classes else None
        rc = self._lib.loco_set_head_outputs(self._h, hp.data_ptr(), lg.data_ptr() if lg is not None else None)
        _lib.check(self._lib, self._h, rc, "loco_set_head_outputs")
        return hp, lg

    def _head_outputs_off(self, with_head: bool):
        """The launch has captured the pointers; no later call (e.g. ``encode_host``) may write through them again."""
        if with_head:
            _lib.check(self._lib, self._h, self._lib.loco_set_head_outputs(self._h, None, None), "loco_set_head_outputs")

    def encode_packed(self, wave: torch.Tensor, n_samples: Sequence[int], return_hidden: bool = False,
                      out: Optional[torch.Tensor] = None, with_head: bool = False):
        """wave: f32[sum(n_samples)] on this device, utterances concatenated without padding.
        Returns pooled f32[B, 768] (and the compact last_hidden_state f32[sum T, 768] if asked).
        ``with_head=True`` (after ``set_head``): returns ``(pooled, head_pooled f32[B,768], logits f32[B,C] or None)``.
        Asynchronous on the current stream."""
        self.finalize()
        if wave.device != self.device or wave.dtype != torch.float32 or not wave.is_contiguous():
            raise _lib.LocoError("encode_packed wants a contiguous float32 waveform tensor on " + str(self.device))
        ns = np.ascontiguousarray(np.asarray(n_samples, dtype=np.int32))
        n = int(ns.shape[0])
        if int(ns.sum()) != wave.numel():
            raise _lib.LocoError(f"sum(n_samples)={int(ns.sum())} does not match the waveform length {wave.numel()}")
        info = self.plan(ns)
        ws = self._get_workspace(info["workspace_bytes"])
        pooled = out if out is not None else torch.empty(n, self.config.hidden_size, dtype=torch.float32, device=self.device)
        hidden = torch.empty(info["total_frames"], self.config.hidden_size, dtype=torch.float32, device=self.device) if return_hidden else None
        head_pooled, logits = self._head_outputs(n, with_head)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = self._lib.loco_encode(self._h, wave.data_ptr(), ns.ctypes.data, n, pooled.data_ptr(),
                                       hidden.data_ptr() if hidden is not None else None, ws.data_ptr(),
                                       ws.numel(), C.c_void_p(stream))
        self._head_outputs_off(with_head)
        _lib.check(self._lib, self._h, rc, "loco_encode")
        self._last_plan = info
        if with_head:
            info = dict(info, head_pooled=head_pooled, logits=logits)
            if return_hidden:
                return pooled, hidden, info
            return pooled, head_pooled, logits
        if return_hidden:
            return pooled, hidden, info
        return pooled

    # ------------------------------------------------------------------ host buffers in, host buffers out
    def encode_host(self, wave_host: torch.Tensor, n_samples: Sequence[int], pooled_host: Optional[torch.Tensor] = None):
        """wave_host: f32[sum(n_samples)] in (ideally pinned) HOST memory; returns pooled f32[B,768] on the host.
        Includes the H2D copy of the waveforms and the D2H copy of the result; synchronous."""
        self.finalize()
        ns = np.ascontiguousarray(np.asarray(n_samples, dtype=np.int32))
        n = int(ns.shape[0])
        if wave_host.device.type != "cpu" or wave_host.dtype != torch.float32 or not wave_host.is_contiguous():
            raise _lib.LocoError("encode_host wants a contiguous float32 CPU tensor")
        if int(ns.sum()) != wave_host.numel():
            raise _lib.LocoError("sum(n_samples) does not match the waveform length")
        need = C.c_size_t()
        rc = self._lib.loco_host_workspace_bytes(self._h, ns.ctypes.data, n, 0, C.byref(need))
        _lib.check(self._lib, self._h, rc, "loco_host_workspace_bytes")
        ws = self._get_workspace(int(need.value))
        if pooled_host is None:
            pooled_host = torch.empty(n, self.config.hidden_size, dtype=torch.float32).pin_memory()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = self._lib.loco_encode_host(self._h, wave_host.data_ptr(), ns.ctypes.data, n, pooled_host.data_ptr(), None,
                                            ws.data_ptr(), ws.numel(), C.c_void_p(stream))
        _lib.check(self._lib, self._h, rc, "loco_encode_host")
        return pooled_host

    def encode_host_pipelined(self, batches, depth: int = 2):
        """Bulk extraction from HOST buffers with the copies hidden behind the compute: ``batches`` yields
        ``(wave_host f32[sum n] (pinned), n_samples[, pooled_host f32[B, 768] (pinned)])``; pooled host tensors are yielded
        in order, one batch late.  Batch i+1's waveforms cross PCIe on a copy stream (into one of ``depth`` staging buffers)
        while batch i is encoded on the current stream; the D2H of the pooled result follows its encode.  Same arithmetic
        as ``encode_host`` / ``loco_encode_host`` (which serialise copy -> encode -> copy per call)."""
        self.finalize()
        dev = self.device
        compute = torch.cuda.current_stream(dev)
        copier = torch.cuda.Stream(dev)
        stage = [None] * depth                     # device staging buffers for the waveforms
        free_ev = [None] * depth                   # recorded on `compute` when the encode that read stage[k] was enqueued
        pending = []                               # (done_event, pooled_host)
        for i, item in enumerate(batches):
            wave_host, n_samples = item[0], item[1]
            pooled_host = item[2] if len(item) > 2 else None
            if wave_host.device.type != "cpu" or wave_host.dtype != torch.float32 or not wave_host.is_contiguous():
                raise _lib.LocoError("encode_host_pipelined wants contiguous float32 CPU tensors")
            k = i % depth
            n = wave_host.numel()
            if stage[k] is None or stage[k].numel() < n:
                if free_ev[k] is not None:
                    free_ev[k].synchronize()
                stage[k] = torch.empty(int(n * 1.1) + 16, dtype=torch.float32, device=dev)
            with torch.cuda.stream(copier):
                if free_ev[k] is not None:
                    copier.wait_event(free_ev[k])           # the encode that last read this staging buffer has finished
                stage[k][:n].copy_(wave_host, non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(copier)
            compute.wait_event(copied)
            pooled_dev = self.encode_packed(stage[k][:n], n_samples)
            free_ev[k] = torch.cuda.Event()
            free_ev[k].record(compute)
            if pooled_host is None:
                pooled_host = torch.empty(pooled_dev.shape, dtype=torch.float32).pin_memory()
            pooled_host.copy_(pooled_dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(compute)
            pending.append((done, pooled_host))
            if len(pending) > 1:
                ev, out = pending.pop(0)
                ev.synchronize()
                yield out
        for ev, out in pending:
            ev.synchronize()
            yield out

    # ------------------------------------------------------------------ text modality (reference :79-93)
    def encode_text_packed(self, tokens: torch.Tensor, n_tokens: Sequence[int], return_hidden: bool = False, with_head: bool = False):
        """tokens: int32[sum(n_tokens)] on this device, texts concatenated without padding.  Each text is encoded alone.
        Returns pooled f32[B, 768] (and the compact last_hidden_state f32[sum n_tokens, 768] if asked); ``with_head`` as in
        ``encode_packed``."""
        self.finalize()
        if tokens.device != self.device or tokens.dtype != torch.int32 or not tokens.is_contiguous():
            raise _lib.LocoError("encode_text_packed wants a contiguous int32 token tensor on " + str(self.device))
        nt = np.ascontiguousarray(np.asarray(n_tokens, dtype=np.int32))
        n = int(nt.shape[0])
        if int(nt.sum()) != tokens.numel():
            raise _lib.LocoError(f"sum(n_tokens)={int(nt.sum())} does not match the token count {tokens.numel()}")
        total, wsb = C.c_int64(), C.c_size_t()
        rows = np.zeros(n, dtype=np.int32)
        rc = self._lib.loco_plan_text(self._h, nt.ctypes.data, n, rows.ctypes.data, C.byref(total), C.byref(wsb))
        _lib.check(self._lib, self._h, rc, "loco_plan_text")
        ws = self._get_workspace(int(wsb.value))
        pooled = torch.empty(n, self.config.hidden_size, dtype=torch.float32, device=self.device)
        hidden = torch.empty(int(total.value), self.config.hidden_size, dtype=torch.float32, device=self.device) if return_hidden else None
        head_pooled, logits = self._head_outputs(n, with_head)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = self._lib.loco_encode_text(self._h, tokens.data_ptr(), nt.ctypes.data, n, pooled.data_ptr(),
                                            hidden.data_ptr() if hidden is not None else None, ws.data_ptr(), ws.numel(),
                                            C.c_void_p(stream))
        self._head_outputs_off(with_head)
        _lib.check(self._lib, self._h, rc, "loco_encode_text")
        info = {"frames": nt.copy(), "rows": rows, "total_frames": int(total.value), "workspace_bytes": int(wsb.value)}
        self._last_plan = info
        if with_head:
            info = dict(info, head_pooled=head_pooled, logits=logits)
            if return_hidden:
                return pooled, hidden, info
            return pooled, head_pooled, logits
        if return_hidden:
            return pooled, hidden, info
        return pooled

    def _call_text(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], return_last_hidden_state: bool):
        """``encoder(texts.input_ids)`` of the reference's text branch (:88): int[B, L_max].  The reference passes no mask,
        so padded positions are real tokens to it; the same happens here.  With a mask each text keeps its own tokens."""
        ids = input_ids.to(self.device)
        if ids.dim() == 1:
            ids = ids[None]
        B, Lmax = ids.shape
        if attention_mask is None:
            lengths = np.full(B, Lmax, dtype=np.int32)
            tok = ids.reshape(-1)
        else:
            am = attention_mask.to(self.device)
            lengths = am.sum(dim=-1).to(torch.int32).cpu().numpy()
            keep = torch.arange(Lmax, device=self.device)[None, :] < torch.as_tensor(lengths, device=self.device)[:, None]
            tok = ids[keep]
        if tok.numel() and (int(tok.min()) < 0 or int(tok.max()) >= self.text_vocab_size()):
            raise IndexError("token id out of range for the text prenet's embedding table")      # what nn.Embedding raises
        tok = tok.to(torch.int32).contiguous()
        if not return_last_hidden_state:
            return LocoEncoderOutput(None, self.encode_text_packed(tok, lengths), torch.as_tensor(lengths, dtype=torch.int64))
        pooled, hidden, info = self.encode_text_packed(tok, lengths, return_hidden=True)
        frames = torch.as_tensor(lengths, dtype=torch.int64)
        tmax = int(frames.max())
        out = torch.zeros(B, tmax, self.config.hidden_size, dtype=torch.float32, device=self.device)
        keep_f = torch.arange(tmax, device=self.device)[None, :] < frames.to(self.device)[:, None]
        out[keep_f] = hidden
        return LocoEncoderOutput(out, pooled, frames)

    def text_vocab_size(self) -> int:
        return int(self._text_vocab)

    # ------------------------------------------------------------------ the reference's call: encoder(**audios)
    def __call__(self, input_values: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                 output_attentions=None, output_hidden_states=None, return_dict=None, return_last_hidden_state: bool = True,
                 **kwargs) -> LocoEncoderOutput:
        """``input_values`` f32[B, L_max] zero-padded, ``attention_mask`` int[B, L_max] (1 = real sample), as
        produced by ``SpeechT5Processor(audio=..., padding="longest")`` (reference :60)."""
        if output_attentions or output_hidden_states:
            raise _lib.LocoError("attention maps / per-layer hidden states are never materialised by the fused kernels")
        if not torch.is_floating_point(input_values):      # token ids: the text branch of the reference scripts
            return self._call_text(input_values, attention_mask, return_last_hidden_state)
        if input_values.dim() == 1:
            input_values = input_values[None]
        iv = input_values.to(self.device, dtype=torch.float32)
        B, Lmax = iv.shape
        if attention_mask is None:
            lengths = np.full(B, Lmax, dtype=np.int32)
            wave = iv.reshape(-1).contiguous()
        else:
            am = attention_mask.to(self.device)
            lengths = am.sum(dim=-1).to(torch.int32).cpu().numpy()   # HF: cumsum(-1)[:, -1]
            keep = torch.arange(Lmax, device=self.device)[None, :] < torch.as_tensor(lengths, device=self.device)[:, None]
            wave = iv[keep].contiguous()
        if not return_last_hidden_state:
            pooled = self.encode_packed(wave, lengths)
            return LocoEncoderOutput(None, pooled, torch.as_tensor(self.plan(lengths)["frames"], dtype=torch.int64))
        pooled, hidden, info = self.encode_packed(wave, lengths, return_hidden=True)
        frames = torch.as_tensor(info["frames"], dtype=torch.int64)
        tmax = int(frames.max())
        out = torch.zeros(B, tmax, self.config.hidden_size, dtype=torch.float32, device=self.device)
        fr = frames.to(self.device)
        keep_f = torch.arange(tmax, device=self.device)[None, :] < fr[:, None]
        out[keep_f] = hidden
        return LocoEncoderOutput(out, pooled, frames)

    forward = __call__

    # ------------------------------------------------------------------ test hooks
    def debug_set(self, name: str, value: int):
        rc = self._lib.loco_debug_set(self._h, name.encode(), int(value))
        _lib.check(self._lib, self._h, rc, "loco_debug_set")

    def debug_buffer(self, name: str) -> torch.Tensor:
        """Copy of a named stage buffer of the last encode as a [rows, cols] bf16 tensor."""
        ptr, rows, cols, dt = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int()
        rc = self._lib.loco_debug_buffer(self._h, name.encode(), C.byref(ptr), C.byref(rows), C.byref(cols), C.byref(dt))
        _lib.check(self._lib, self._h, rc, "loco_debug_buffer")
        off = ptr.value - self._workspace.data_ptr()
        nbytes = rows.value * cols.value * 2
        return self._workspace[off:off + nbytes].view(torch.bfloat16).view(rows.value, cols.value).clone()

    def debug_gemm(self, a, w, bias=None, residual=None, epilogue=_lib.EPI_BIAS, impl=0, lda=None, m=None):
        """C = epi(A W^T): unit-test entry for the GEMM kernels (a: bf16 [rows, K] or flat with lda)."""
        k = w.shape[1]
        n = w.shape[0]
        lda = lda if lda is not None else a.shape[-1]
        m = m if m is not None else a.shape[0]
        rows_alloc = (a.numel() - k) // lda + 1
        c = torch.empty(m, n, dtype=torch.bfloat16, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = self._lib.loco_debug_gemm(self._h, impl, a.data_ptr(), lda, rows_alloc, w.data_ptr(), c.data_ptr(),
                                           bias.data_ptr() if bias is not None else None,
                                           residual.data_ptr() if residual is not None else None, m, n, k, epilogue,
                                           C.c_void_p(stream))
        _lib.check(self._lib, self._h, rc, "loco_debug_gemm")
        return c

    def debug_gemm_ln(self, a, w, epilogue, bias=None, residual=None, stats_in=None, c1=None, gamma=None, want_stats=False):
        """Unit-test entry for the deferred-LayerNorm epilogues of the CTA-pair GEMM (csrc/internal.h GemmEpilogue 3..6).
        Returns C bf16 [M, N] (and the row statistics f32 [M, 6, 2] when asked)."""
        m, k = a.shape
        n = w.shape[0]
        c = torch.empty(m, n, dtype=torch.bfloat16, device=self.device)
        stats = torch.zeros(m, 6, 2, dtype=torch.float32, device=self.device) if want_stats else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = self._lib.loco_debug_gemm_ln(self._h, a.data_ptr(), w.data_ptr(), c.data_ptr(), ptr(bias), ptr(residual), m, n, k,
                                              epilogue, ptr(stats_in), ptr(c1), ptr(gamma), ptr(stats), C.c_void_p(stream))
        _lib.check(self._lib, self._h, rc, "loco_debug_gemm_ln")
        return (c, stats) if want_stats else c

    PROFILE_CATEGORIES = ("gemm", "attention", "pos_conv", "frontend", "rowops")

    def profile_enable(self, on: bool = True):
        rc = self._lib.loco_profile_enable(self._h, 1 if on else 0)
        _lib.check(self._lib, self._h, rc, "loco_profile_enable")

    def profile_collect(self):
        """{category: (milliseconds, launches)} since the last collect (synchronises the device)."""
        ms = (C.c_double * 5)()
        n = (C.c_int64 * 5)()
        rc = self._lib.loco_profile_collect(self._h, 5, ms, n)
        _lib.check(self._lib, self._h, rc, "loco_profile_collect")
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.PROFILE_CATEGORIES)}

    @property
    def launch_count(self) -> int:
        return int(self._lib.loco_launch_count(self._h))
