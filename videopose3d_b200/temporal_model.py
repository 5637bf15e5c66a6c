"""Drop-in replacements for VideoPose3D's temporal-convolution models.

Mirrors ``common/model.py`` of the reference (TemporalModelBase :10-77, TemporalModel :79-138,
TemporalModelOptimized1f :140-197): same constructor signatures, attribute names and
``state_dict`` keys/shapes, so ``run.py`` and published checkpoints work unchanged.  Parameters
live in ordinary ``nn.Conv1d`` / ``nn.BatchNorm1d`` containers (never called); ``forward`` hands raw
device pointers to the sm_100a kernels behind the C ABI (``include/vp3d_b200.h``).

There is no PyTorch/CPU execution path here: without the CUDA library or with CPU tensors
``forward`` raises.
"""
import os

import torch
import torch.nn as nn

from . import _capi

_PRECISIONS = {"bf16": _capi.VP3D_PRECISION_BF16, "bf16x3": _capi.VP3D_PRECISION_BF16X3}


def _default_precision():
    p = os.environ.get("VP3D_PRECISION", "bf16")
    if p not in _PRECISIONS:
        raise ValueError(f"VP3D_PRECISION must be one of {sorted(_PRECISIONS)}, got {p!r}")
    return p


class TemporalModelBase(nn.Module):
    """
    Do not instantiate this class.  (reference: common/model.py:10-77)
    """

    _variant = None  # set by subclasses

    def __init__(self, num_joints_in, in_features, num_joints_out,
                 filter_widths, causal, dropout, channels):
        super().__init__()

        # Validate input (model.py:20-21)
        for fw in filter_widths:
            assert fw % 2 != 0, 'Only odd filter widths are supported'

        self.num_joints_in = num_joints_in
        self.in_features = in_features
        self.num_joints_out = num_joints_out
        self.filter_widths = filter_widths

        self.drop = nn.Dropout(dropout)
        self.relu = nn.ReLU(inplace=True)

        self.pad = [filter_widths[0] // 2]
        self.expand_bn = nn.BatchNorm1d(channels, momentum=0.1)
        self.shrink = nn.Conv1d(channels, num_joints_out * 3, 1)

        # engine state (not part of the state_dict)
        self._channels = channels
        self._causal = bool(causal)
        self._dense = False
        self._precision = _default_precision()
        self._plan = None
        self._plan_key = None
        self._packed_versions = None
        self._workspace = None

    def _build_layers(self, strided):
        """Create the parameter containers with the reference's names, shapes and default init.

        Per block i >= 1 with width w_i and dilation d_i = prod(w_0..w_{i-1}) (model.py:107-121,
        172-184): pad_i = (w_i - 1) * d_i // 2; the first conv is Conv1d(C, C, w_i, dilation=d_i)
        for the dilated model (kernel 2*pad_i+1, dilation 1 when dense) or Conv1d(C, C, w_i,
        stride=w_i) for the strided one; the second is a 1x1 conv; each is followed by a
        BatchNorm1d(momentum=0.1).  causal_shift is in frames for the dilated model and in strided
        units for the strided one.
        """
        fw, C = self.filter_widths, self._channels
        c_in = self.num_joints_in * self.in_features
        half0 = fw[0] // 2
        if strided:
            self.expand_conv = nn.Conv1d(c_in, C, fw[0], stride=fw[0], bias=False)
        else:
            self.expand_conv = nn.Conv1d(c_in, C, fw[0], bias=False)
        self.causal_shift = [half0 if self._causal else 0]
        convs, bns = [], []
        dilation = fw[0]
        for w in fw[1:]:
            pad = (w - 1) * dilation // 2
            self.pad.append(pad)
            if strided:
                self.causal_shift.append(w // 2 if self._causal else 0)
                convs.append(nn.Conv1d(C, C, w, stride=w, bias=False))
            else:
                self.causal_shift.append((w // 2) * dilation if self._causal else 0)
                if self._dense:
                    convs.append(nn.Conv1d(C, C, 2 * pad + 1, dilation=1, bias=False))
                else:
                    convs.append(nn.Conv1d(C, C, w, dilation=dilation, bias=False))
            bns.append(nn.BatchNorm1d(C, momentum=0.1))
            convs.append(nn.Conv1d(C, C, 1, dilation=1, bias=False))
            bns.append(nn.BatchNorm1d(C, momentum=0.1))
            dilation *= w
        self.layers_conv = nn.ModuleList(convs)
        self.layers_bn = nn.ModuleList(bns)

    # ------------------------------------------------------------------ reference API
    def set_bn_momentum(self, momentum):
        # model.py:36-39 — read at call time by the training kernels, never cached
        self.expand_bn.momentum = momentum
        for bn in self.layers_bn:
            bn.momentum = momentum

    def receptive_field(self):
        """
        Return the total receptive field of this model as # of frames.  (model.py:41-48)
        """
        frames = 0
        for f in self.pad:
            frames += f
        return 1 + 2 * frames

    def total_causal_shift(self):
        """
        Return the asymmetric offset for sequence padding.  (model.py:50-61)
        """
        frames = self.causal_shift[0]
        next_dilation = self.filter_widths[0]
        for i in range(1, len(self.filter_widths)):
            frames += self.causal_shift[i] * next_dilation
            next_dilation *= self.filter_widths[i]
        return frames

    # ------------------------------------------------------------------ engine controls
    def set_precision(self, precision):
        """'bf16' (fast path) or 'bf16x3' (split-bf16, fp32-faithful).  Not in the reference."""
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        if precision != self._precision:
            self._precision = precision
            self._release_plan()
        return self

    @property
    def precision(self):
        return self._precision

    def _release_plan(self):
        if self._plan is not None:
            try:
                _capi.load().vp3d_plan_destroy(self._plan)
            except Exception:  # pragma: no cover - interpreter shutdown
                pass
        self._plan = None
        self._plan_key = None
        self._packed_versions = None

    def __del__(self):
        try:
            self._release_plan()
        except Exception:  # pragma: no cover
            pass

    def _config(self):
        cfg = _capi.Config()
        cfg.num_joints_in = self.num_joints_in
        cfg.in_features = self.in_features
        cfg.num_joints_out = self.num_joints_out
        cfg.num_widths = len(self.filter_widths)
        if cfg.num_widths > _capi.VP3D_MAX_WIDTHS:
            raise NotImplementedError(f"at most {_capi.VP3D_MAX_WIDTHS} filter widths are supported")
        for i, w in enumerate(self.filter_widths):
            cfg.filter_widths[i] = int(w)
        cfg.causal = int(self._causal)
        cfg.channels = self._channels
        cfg.dense = int(self._dense)
        cfg.variant = self._variant
        cfg.precision = _PRECISIONS[self._precision]
        return cfg

    def _get_plan(self, device):
        key = (device.index, self._precision)
        if self._plan is None or self._plan_key != key:
            self._release_plan()
            lib = _capi.load()
            handle = _capi.ctypes.c_void_p()
            cfg = self._config()
            with torch.cuda.device(device):
                _capi.check(lib.vp3d_plan_create(_capi.ctypes.byref(cfg), _capi.ctypes.byref(handle)),
                            "vp3d_plan_create")
            self._plan = handle
            self._plan_key = key
            self._packed_versions = None
        return self._plan

    def _param_tensors(self):
        """All fp32 tensors of the state_dict in the order of ``vp3d_weights``."""
        conv = [self.expand_conv.weight] + [c.weight for c in self.layers_conv] + \
               [self.shrink.weight]
        bn = []
        for m in [self.expand_bn] + list(self.layers_bn):
            bn += [m.weight, m.bias, m.running_mean, m.running_var]
        bn.append(self.shrink.bias)
        return conv, bn

    def _sync_weights(self, plan, stream):
        conv, bn = self._param_tensors()
        versions = (tuple((t.data_ptr(), t._version) for t in conv),
                    tuple((t.data_ptr(), t._version) for t in bn))
        what = 0
        if self._packed_versions is None or self._packed_versions[0] != versions[0]:
            what |= _capi.VP3D_PACK_CONV
        if self._packed_versions is None or self._packed_versions[1] != versions[1]:
            what |= _capi.VP3D_PACK_BN_EVAL
        if not what:
            return
        for t in conv + bn:
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("parameters must be contiguous float32 tensors")
        w = _capi.Weights()
        w.expand_conv_weight = self.expand_conv.weight.data_ptr()
        for k, t in enumerate((self.expand_bn.weight, self.expand_bn.bias,
                               self.expand_bn.running_mean, self.expand_bn.running_var)):
            w.expand_bn[k] = t.data_ptr()
        for i, c in enumerate(self.layers_conv):
            w.layers_conv_weight[i] = c.weight.data_ptr()
        for i, m in enumerate(self.layers_bn):
            for k, t in enumerate((m.weight, m.bias, m.running_mean, m.running_var)):
                w.layers_bn[i][k] = t.data_ptr()
        w.shrink_weight = self.shrink.weight.data_ptr()
        w.shrink_bias = self.shrink.bias.data_ptr()
        _capi.check(_capi.load().vp3d_set_weights(plan, _capi.ctypes.byref(w), what, stream),
                    "vp3d_set_weights")
        self._packed_versions = versions

    def _get_workspace(self, nbytes, device):
        ws = self._workspace
        if ws is None or ws.device != device or ws.numel() < nbytes:
            self._workspace = None
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._workspace = ws
        return ws

    # ------------------------------------------------------------------ forward
    def forward(self, x):
        assert len(x.shape) == 4
        assert x.shape[-2] == self.num_joints_in
        assert x.shape[-1] == self.in_features

        if not x.is_cuda:
            raise RuntimeError("videopose3d_b200 models run on CUDA (sm_100a) tensors only; "
                               "there is no CPU fallback")
        if x.dtype != torch.float32:
            raise TypeError(f"expected a float32 input, got {x.dtype}")
        if self.expand_conv.weight.device != x.device:
            raise RuntimeError("input and parameters are on different devices")
        if self.training:
            return self._forward_train(x)
        return self._forward_eval(x)

    def _forward_eval(self, x):
        lib = _capi.load()
        x = x.contiguous()
        device = x.device
        N, T = int(x.shape[0]), int(x.shape[1])
        with torch.cuda.device(device):
            plan = self._get_plan(device)
            stream = torch.cuda.current_stream(device).cuda_stream
            self._sync_weights(plan, stream)
            t_out = lib.vp3d_output_frames(plan, T)
            if t_out < 1:
                raise ValueError(f"input of {T} frames is shorter than the receptive field "
                                 f"({self.receptive_field()})")
            nbytes = lib.vp3d_workspace_bytes(plan, N, T)
            ws = self._get_workspace(nbytes, device)
            y = torch.empty((N, t_out, self.num_joints_out, 3), dtype=torch.float32, device=device)
            _capi.check(lib.vp3d_forward_eval(plan, x.data_ptr(), y.data_ptr(), N, T, ws.data_ptr(),
                                              ws.numel(), stream), "vp3d_forward_eval")
        return y

    def _forward_train(self, x):
        raise NotImplementedError("training-mode forward/backward kernels are not built yet "
                                  "(round-1 scope: eval forward); call .eval() first")

    def forward_host(self, x_host, out=None):
        """Eval forward from a HOST float32 tensor/array (pinned for full PCIe speed): copies the
        batch to the device, runs the kernels and copies the result back (the .cuda()/.cpu() round
        trip of run.py:663-672 in one call).  Used by bench.py for the end-to-end number."""
        lib = _capi.load()
        if self.training:
            raise RuntimeError("forward_host is an eval-mode call")
        x_host = torch.as_tensor(x_host)
        assert x_host.dim() == 4 and x_host.shape[-2] == self.num_joints_in \
            and x_host.shape[-1] == self.in_features
        if x_host.is_cuda or x_host.dtype != torch.float32 or not x_host.is_contiguous():
            raise ValueError("forward_host expects a contiguous float32 CPU tensor")
        device = self.expand_conv.weight.device
        if device.type != "cuda":
            raise RuntimeError("module parameters must be on a CUDA device")
        N, T = int(x_host.shape[0]), int(x_host.shape[1])
        with torch.cuda.device(device):
            plan = self._get_plan(device)
            stream = torch.cuda.current_stream(device)
            self._sync_weights(plan, stream.cuda_stream)
            stream.synchronize()  # packed weights are read by the plan's own stream
            t_out = lib.vp3d_output_frames(plan, T)
            if t_out < 1:
                raise ValueError("input shorter than the receptive field")
            if out is None:
                out = torch.empty((N, t_out, self.num_joints_out, 3), dtype=torch.float32,
                                  pin_memory=True)
            _capi.check(lib.vp3d_forward_eval_host(plan, x_host.data_ptr(), out.data_ptr(), N, T),
                        "vp3d_forward_eval_host")
        return out

    def last_launch_count(self):
        return 0 if self._plan is None else _capi.load().vp3d_last_launch_count(self._plan)


class TemporalModel(TemporalModelBase):
    """Dilated-convolution model, usable for every use-case (reference: common/model.py:79-138).

    Constructor signature identical to model.py:85-86:
    num_joints_in, in_features, num_joints_out, filter_widths, causal=False, dropout=0.25,
    channels=1024, dense=False (dense = ablation with regular convolutions of width 2*pad+1).
    """

    _variant = _capi.VP3D_VARIANT_DILATED

    def __init__(self, num_joints_in, in_features, num_joints_out,
                 filter_widths, causal=False, dropout=0.25, channels=1024, dense=False):
        super().__init__(num_joints_in, in_features, num_joints_out, filter_widths, causal, dropout,
                         channels)
        self._dense = bool(dense)
        self._build_layers(strided=False)


class TemporalModelOptimized1f(TemporalModelBase):
    """Strided model for single-frame batches: input length == receptive field, one output frame
    (reference: common/model.py:140-197).  Same parameters as TemporalModel, interchangeable
    weights; constructor signature identical to model.py:151-152.
    """

    _variant = _capi.VP3D_VARIANT_STRIDED

    def __init__(self, num_joints_in, in_features, num_joints_out,
                 filter_widths, causal=False, dropout=0.25, channels=1024):
        super().__init__(num_joints_in, in_features, num_joints_out, filter_widths, causal, dropout,
                         channels)
        self._build_layers(strided=True)
