"""``attention_aspp_unet`` -- the module name the reference imports and never ships (model_attention_aspp.py:6).

Drop-in for the reference's ``AttentionASPPUNet`` ``nn.Module``
(attention_aspp_unet_pipeline_stage.py:111-127; ablation twin test_ablation.py:168-218):

* same constructor (both spellings the reference uses: ``in_channels/base_c`` and ``in_ch/base``; the ablation
  flags ``use_att / use_aspp / att_depth`` select the ablation twin),
* same ``state_dict()`` keys, shapes and order (real ``nn.Parameter`` / buffers, so checkpoints written by the
  reference's ``torch.save(model.state_dict())`` load with ``load_state_dict(sd, strict=...)``),
* ``forward(x[B,1,H,W] fp32 cuda) -> logits[B,1,H,W] fp32`` (ablation twin: ``(logits, [psi3, psi2])``).

The forward itself is NOT PyTorch: it is one call into ``libaau.so`` (hand-written sm_100a kernels, include/aau.h).
There is no CPU path and no eager fallback; a missing library or a non-CUDA input raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Tuple

import torch
import torch.nn as nn

try:                                    # imported as a top-level module (directory on sys.path) ...
    import _capi
except ImportError:                     # ... or as part of a package
    from . import _capi                 # type: ignore

__all__ = ["AttentionASPPUNet", "param_spec"]

_UNSET = object()


# ------------------------------------------------------------------------------------------------------------
# parameter table (the state_dict contract)
# ------------------------------------------------------------------------------------------------------------
def _bn(prefix: str, n: int):
    return [(f"{prefix}.weight", (n,), "ones", True), (f"{prefix}.bias", (n,), "zeros", True),
            (f"{prefix}.running_mean", (n,), "zeros", False), (f"{prefix}.running_var", (n,), "ones", False),
            (f"{prefix}.num_batches_tracked", (), "count", False)]


def _cbr(prefix: str, cin: int, cout: int):
    return [(f"{prefix}.block.0.weight", (cout, cin, 3, 3), "conv", True)] + _bn(f"{prefix}.block.1", cout)


def gate_levels(variant: str, use_att: bool, att_depth: int) -> Tuple[int, ...]:
    if variant == "pipeline":
        return (4, 3, 2)
    return tuple(l for l in (4, 3) if use_att and att_depth >= l)


def f_int(variant: str, base_c: int, level: int) -> int:
    out_c = base_c << (level - 1)
    return out_c // 2 if variant == "pipeline" else max(8, out_c // 4)


def param_spec(base_c: int, variant: str = "pipeline", use_att: bool = True, use_aspp: bool = True, att_depth: int = 4,
               in_channels: int = 1, num_classes: int = 1) -> List[Tuple[str, Tuple[int, ...], str, bool]]:
    """``(key, shape, init, is_parameter)`` for every ``state_dict`` entry, in the reference's order."""
    c = base_c
    width = [in_channels, c, 2 * c, 4 * c, 8 * c]
    rows: List[Tuple[str, Tuple[int, ...], str, bool]] = []
    for lvl in (1, 2, 3, 4):
        rows += _cbr(f"d{lvl}.0", width[lvl - 1], width[lvl]) + _cbr(f"d{lvl}.1", width[lvl], width[lvl])
    if variant == "pipeline" or use_aspp:
        ic, oc = 8 * c, 16 * c
        rows += [("bridge.blocks.0.0.weight", (oc, ic, 1, 1), "conv", True)] + _bn("bridge.blocks.0.1", oc)
        for i in (1, 2, 3):
            rows += [(f"bridge.blocks.{i}.0.weight", (oc, ic, 3, 3), "conv", True)] + _bn(f"bridge.blocks.{i}.1", oc)
        rows += [("bridge.pool.1.weight", (oc, ic, 1, 1), "conv", True)] + _bn("bridge.pool.2", oc)
        rows += [("bridge.project.0.weight", (oc, 5 * oc, 1, 1), "conv", True)] + _bn("bridge.project.1", oc)
    else:
        rows += _cbr("bridge.0", 8 * c, 16 * c)
    gates = gate_levels(variant, use_att, att_depth)
    for lvl in (4, 3, 2, 1):
        oc = c << (lvl - 1)
        ic = 2 * oc
        u = f"u{lvl}"
        rows += [(f"{u}.up.weight", (ic, oc, 2, 2), "conv", True), (f"{u}.up.bias", (oc,), f"bias:{oc * 4}", True)]
        if lvl in gates:
            fi = f_int(variant, c, lvl)
            if variant == "pipeline":
                rows += [(f"{u}.att.Wg.0.weight", (fi, oc, 1, 1), "conv", True)] + _bn(f"{u}.att.Wg.1", fi)
                rows += [(f"{u}.att.Wx.0.weight", (fi, oc, 1, 1), "conv", True)] + _bn(f"{u}.att.Wx.1", fi)
                rows += [(f"{u}.att.psi.0.weight", (1, fi, 1, 1), "conv", True)] + _bn(f"{u}.att.psi.1", 1)
            else:
                rows += [(f"{u}.att.Wg.weight", (fi, oc, 1, 1), "conv", True), (f"{u}.att.Wx.weight", (fi, oc, 1, 1), "conv", True),
                         (f"{u}.att.psi.1.weight", (1, fi, 1, 1), "conv", True), (f"{u}.att.psi.1.bias", (1,), f"bias:{fi}", True)]
        rows += _cbr(f"{u}.conv.0", ic, oc) + _cbr(f"{u}.conv.1", oc, oc)
    rows += [("out_conv.weight", (num_classes, c, 1, 1), "conv", True), ("out_conv.bias", (num_classes,), f"bias:{c}", True)]
    return rows


class _Node(nn.Module):
    """Anonymous container: gives parameters the dotted names of the reference's nested modules."""


def _attach(root: nn.Module, dotted: str, tensor: torch.Tensor, is_param: bool) -> None:
    *path, leaf = dotted.split(".")
    node = root
    for part in path:
        if part not in node._modules:
            node.add_module(part, _Node())
        node = node._modules[part]
    if is_param:
        node.register_parameter(leaf, nn.Parameter(tensor))
    else:
        node.register_buffer(leaf, tensor)


def _init_tensor(shape, init: str) -> torch.Tensor:
    if init == "ones":
        return torch.ones(shape)
    if init == "zeros":
        return torch.zeros(shape)
    if init == "count":
        return torch.zeros(shape, dtype=torch.long)
    if init == "conv":                      # nn.Conv2d / ConvTranspose2d default: U(+-1/sqrt(fan_in)), fan_in = shape[1]*kh*kw
        bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
        return torch.empty(shape).uniform_(-bound, bound)
    if init.startswith("bias:"):
        bound = 1.0 / math.sqrt(int(init[5:]))
        return torch.empty(shape).uniform_(-bound, bound)
    raise ValueError(init)


# ------------------------------------------------------------------------------------------------------------
# the module
# ------------------------------------------------------------------------------------------------------------
class AttentionASPPUNet(nn.Module):
    """B200-native AttentionASPPUNet (inference only).

    ``AttentionASPPUNet(in_channels=1, num_classes=1, base_c=32)`` is the canonical model
    (attention_aspp_unet_pipeline_stage.py:112); ``AttentionASPPUNet(in_ch=1, num_classes=1, base=16)`` is how the
    ROI wrapper spells it (model_attention_aspp.py:36); passing any of ``use_att / use_aspp / att_depth`` (or
    ``variant="ablation"``) builds the ablation twin (test_ablation.py:169-177) whose forward returns
    ``(logits, [psi3, psi2])``.  ``act_dtype`` is the 16-bit storage type of activations and packed weights inside
    the engine -- "fp16" (default: 11 significant bits, the only 16-bit type that meets the 2e-2 / 99.9 % parity bar on
    BN-calibrated weights) or "bf16" (8 bits; same kernels, same speed); the tensor cores accumulate
    in fp32 and all epilogue math is fp32 in both.  Values beyond the fp16 range saturate to +-65504 instead of becoming inf.
    """

    def __init__(self, in_channels: int = 1, num_classes: int = 1, base_c: int = 32, use_att=_UNSET, use_aspp=_UNSET,
                 att_depth=_UNSET, *, in_ch=None, base=None, variant: str | None = None, act_dtype: str = "fp16"):
        super().__init__()
        if in_ch is not None:
            in_channels = in_ch
        if base is not None:
            base_c = base
        ablation_args = any(v is not _UNSET for v in (use_att, use_aspp, att_depth))
        if variant is None:
            variant = "ablation" if ablation_args else "pipeline"
        if variant not in ("pipeline", "ablation"):
            raise ValueError("variant must be 'pipeline' or 'ablation'")
        if variant == "pipeline" and ablation_args:
            raise ValueError("use_att / use_aspp / att_depth belong to the ablation twin")
        if in_channels != 1 or num_classes != 1:
            raise ValueError("the B200 engine implements the reference's only configuration: in_channels=1, num_classes=1")
        if base_c < 16 or base_c % 16 or base_c > 64:
            raise ValueError("base_c must be 16, 32, 48 or 64 (the reference uses 16, 32 and 48): the fused gate epilogue holds all "
                             "F_int = 4 * base_c channels of the deepest gate in one 256-column accumulator tile")
        if act_dtype not in ("bf16", "fp16"):
            raise ValueError("act_dtype must be 'bf16' or 'fp16'")
        self.in_channels, self.num_classes, self.base_c = in_channels, num_classes, base_c
        self.variant = variant
        self.use_att = True if use_att is _UNSET else bool(use_att)
        self.use_aspp = True if use_aspp is _UNSET else bool(use_aspp)
        self.att_depth = 4 if att_depth is _UNSET else int(att_depth)
        self.act_dtype = act_dtype
        self._spec = param_spec(base_c, variant, self.use_att, self.use_aspp, self.att_depth, in_channels, num_classes)
        for key, shape, init, is_param in self._spec:
            _attach(self, key, _init_tensor(shape, init), is_param)
        self._handle = None
        self._handle_device = None
        self._weights_version = None
        self._workspaces: Dict[Tuple[int, int, int], torch.Tensor] = {}
        self._last_ws = None

    # ---- engine lifetime -------------------------------------------------------------------------------
    def _config(self) -> _capi.AauConfig:
        return _capi.AauConfig(self.in_channels, self.num_classes, self.base_c,
                               _capi.AAU_VARIANT_PIPELINE if self.variant == "pipeline" else _capi.AAU_VARIANT_ABLATION,
                               int(self.use_att), int(self.use_aspp), self.att_depth,
                               _capi.AAU_ACT_BF16 if self.act_dtype == "bf16" else _capi.AAU_ACT_FP16)

    def _release(self):
        if getattr(self, "_handle", None) is not None:
            try:
                _capi.lib().aau_destroy(self._handle)
            except Exception:
                pass
            self._handle = None
            self._workspaces = {}
            self._weights_version = None

    def __del__(self):
        try:
            self._release()
        except Exception:                                # interpreter shutdown: torch's module machinery may already be gone
            pass

    def _float_entries(self):
        for key, tensor in self.state_dict(keep_vars=True).items():
            if tensor.dtype.is_floating_point:
                yield key, tensor

    def _version(self):
        return tuple((t.data_ptr(), t._version) for _, t in self._float_entries())

    def _ensure_engine(self, device: torch.device):
        L = _capi.lib()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is None or self._handle_device != idx:
            self._release()
            h = C.c_void_p()
            cfg = self._config()
            st = L.aau_create(C.byref(cfg), idx, C.byref(h))
            if st != 0:
                raise _capi.AauError(f"aau_create failed ({st}): {L.aau_last_error(None).decode()}")
            self._handle, self._handle_device = h, idx
        ver = self._version()
        if ver != self._weights_version:
            self.refresh_weights()
            self._weights_version = ver

    def refresh_weights(self):
        """Re-fold BatchNorm and re-pack the weights inside the engine from the module's current parameters."""
        L = _capi.lib()
        if self._handle is None:
            return
        for key, tensor in self._float_entries():
            host = tensor.detach().to("cpu", torch.float32).contiguous()
            _capi.check(self._handle, L.aau_load_tensor(self._handle, key.encode(), host.data_ptr(), host.numel()), f"load {key}")
        _capi.check(self._handle, L.aau_commit_weights(self._handle), "aau_commit_weights")
        self._workspaces = {}

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        # the reference renames legacy gate keys before loading (attention_aspp_unet_pipeline_stage.py:134-141)
        if isinstance(state_dict, dict) and "state_dict" in state_dict and not any(k.startswith("d1.") for k in state_dict):
            state_dict = state_dict["state_dict"]                              # test_ablation.py:224-225
        renamed = {k.replace(".W_g.", ".Wg.").replace(".W_x.", ".Wx."): v for k, v in state_dict.items()}
        out = super().load_state_dict(renamed, strict=strict, assign=assign)
        self._weights_version = None
        return out

    def _workspace(self, B: int, H: int, W: int, device) -> torch.Tensor:
        key = (B, H, W)
        ws = self._workspaces.get(key)
        if ws is None:
            need = _capi.lib().aau_workspace_bytes(self._handle, B, H, W)
            if need == 0:
                raise _capi.AauError(f"unsupported input size {B}x{H}x{W} (H, W must be >= 16)")
            if len(self._workspaces) >= 4:
                self._workspaces.pop(next(iter(self._workspaces)))
            ws = torch.empty(need + 256, dtype=torch.uint8, device=device)
            self._workspaces[key] = ws
        return ws

    # ---- forward ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor, out: torch.Tensor | None = None):
        """``x``: CUDA ``float32 [B,1,H,W]`` in [0,1] (the reference contract) or ``uint8 [B,H,W]`` / ``[B,1,H,W]``
        (normalised on the device as ``u8/255``).  ``out`` (optional) is a contiguous CUDA ``float32 [B,1,H,W]``
        buffer to receive the logits instead of a fresh allocation."""
        if self.training:
            raise RuntimeError("this is an inference engine: call .eval() first (the reference does, model_attention_aspp.py:39)")
        if not x.is_cuda:
            raise RuntimeError("AttentionASPPUNet (B200 engine) needs a CUDA tensor; there is no CPU fallback")
        if x.dtype == torch.uint8:
            if x.dim() == 4:
                x = x[:, 0]
            x_kind = _capi.AAU_X_U8
        else:
            if x.dim() != 4 or x.shape[1] != 1:
                raise ValueError(f"expected [B,1,H,W], got {tuple(x.shape)}")
            x = x.float()
            x_kind = _capi.AAU_X_F32
        x = x.contiguous()
        B, H, W = x.shape[0], x.shape[-2], x.shape[-1]
        self._ensure_engine(x.device)
        L = _capi.lib()
        with torch.cuda.device(x.device):
            ws = self._workspace(B, H, W, x.device)
            self._last_ws = ws                                       # debug_tensor reads the workspace of the LAST forward
            ws_ptr = (ws.data_ptr() + 255) & ~255
            if out is not None:
                if out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != B * H * W or out.device != x.device:
                    raise ValueError("out must be a contiguous CUDA float32 tensor with B*H*W elements on x's device")
                logits = out.view(B, 1, H, W)
            else:
                logits = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
            psi3 = psi2 = None
            p3 = p2 = None
            if self.variant == "ablation":
                gates = gate_levels(self.variant, self.use_att, self.att_depth)
                if 4 in gates:
                    psi3 = torch.empty((B, 1, H // 8, W // 8), dtype=torch.float32, device=x.device)
                    p3 = psi3.data_ptr()
                if 3 in gates:
                    psi2 = torch.empty((B, 1, H // 4, W // 4), dtype=torch.float32, device=x.device)
                    p2 = psi2.data_ptr()
            st = L.aau_forward(self._handle, x.data_ptr(), x_kind, B, H, W, logits.data_ptr(), p3, p2, ws_ptr,
                               ws.numel() - (ws_ptr - ws.data_ptr()), torch.cuda.current_stream(x.device).cuda_stream)
            _capi.check(self._handle, st, "aau_forward")
        if self.variant == "pipeline":
            return logits
        zero = lambda: torch.zeros(1, 1, 1, 1, device=x.device)   # noqa: E731  (test_ablation.py:147)
        return logits, [psi3 if psi3 is not None else zero(), psi2 if psi2 is not None else zero()]

    # ---- measurement / debug aids -----------------------------------------------------------------------
    def engine_handle(self):
        return self._handle

    def num_launches(self) -> int:
        return _capi.lib().aau_num_launches(self._handle) if self._handle is not None else 0

    def last_forward_was_graph(self) -> bool:
        """True when the last forward went out as one CUDA-graph launch (small batches; ``set_option("graph", 0 / 1 / -1)``)."""
        return bool(_capi.lib().aau_last_forward_was_graph(self._handle)) if self._handle is not None else False

    def op_profile(self):
        """Per-launch records of the last forward: ``[{layer, kernel, ms, flops, bytes}]`` (``ms`` is -1 unless
        ``set_option("profile", 1)`` was active).  Synchronises."""
        L = _capi.lib()
        rows = []
        for i in range(L.aau_num_ops(self._handle)):
            layer, kern = C.c_char_p(), C.c_char_p()
            ms, fl, by = C.c_float(), C.c_double(), C.c_double()
            _capi.check(self._handle, L.aau_op_profile(self._handle, i, C.byref(layer), C.byref(kern), C.byref(ms), C.byref(fl), C.byref(by)),
                        "aau_op_profile")
            rows.append({"layer": layer.value.decode(), "kernel": kern.value.decode(), "ms": ms.value, "flops": fl.value, "bytes": by.value})
        return rows

    def set_option(self, name: str, value: int):
        if self._handle is None:
            raise RuntimeError("engine not created yet (run a forward or call .prepare(device))")
        _capi.check(self._handle, _capi.lib().aau_set_option(self._handle, name.encode(), value), "aau_set_option")

    def prepare(self, device="cuda"):
        """Create the engine and upload the weights without running a forward."""
        self._ensure_engine(torch.device(device))
        return self

    def check_device(self):
        """Synchronise and raise if any kernel reported a pipeline fault."""
        if self._handle is not None:
            _capi.check(self._handle, _capi.lib().aau_device_fault(self._handle), "device check")

    def debug_tensor(self, name: str) -> torch.Tensor:
        """fp32 NCHW copy of a named intermediate of the last forward (layer-by-layer parity tests)."""
        L = _capi.lib()
        ptr = C.c_void_p()
        v = [C.c_int() for _ in range(6)]
        _capi.check(self._handle, L.aau_debug_tensor(self._handle, name.encode(), C.byref(ptr), *[C.byref(i) for i in v]), "debug_tensor")
        B, H, W, Cc, ld, choff = [i.value for i in v]
        ws = self._last_ws
        off = ptr.value - ws.data_ptr()
        dt = torch.bfloat16 if self.act_dtype == "bf16" else torch.float16
        flat = ws[off: off + B * H * W * ld * 2].view(dt).view(B, H, W, ld)
        return flat[..., choff:choff + Cc].permute(0, 3, 1, 2).float().contiguous()
