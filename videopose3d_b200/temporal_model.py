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
import weakref

import torch
import torch.nn as nn

from . import _capi

_PRECISIONS = {"bf16": _capi.VP3D_PRECISION_BF16, "bf16x3": _capi.VP3D_PRECISION_BF16X3,
               "mixed": _capi.VP3D_PRECISION_MIXED, "fp16": _capi.VP3D_PRECISION_FP16}


# id(parameter) -> weakref(owning model): lets optim.FusedAdam find the model whose packed bf16
# weights it can refresh while it updates the parameter (no API change for run.py, which builds the
# optimizer from `model.parameters()` alone)
_PARAM_OWNER = {}


def owner_of(param):
    """The live TemporalModel* that owns `param`, or None."""
    ref = _PARAM_OWNER.get(id(param))
    m = ref() if ref is not None else None
    if m is None or not any(param is q for q in m.parameters()):
        return None
    return m


class _PlanStore(dict):
    """(device index, precision) -> plan handle.  Owns the handles: they are destroyed exactly once,
    when the store itself is collected (weakref.finalize), never by a module that merely shares or
    copies the reference.  Copies of a module start with an empty store of their own."""

    def __init__(self):
        super().__init__()
        self._handles = []
        self._finalizer = weakref.finalize(self, _PlanStore._destroy, self._handles)

    def add(self, key, handle):
        self[key] = handle
        self._handles.append(handle)

    def __deepcopy__(self, memo):
        return _PlanStore()

    def __copy__(self):
        return _PlanStore()

    def __reduce__(self):
        return (_PlanStore, ())

    @staticmethod
    def _destroy(handles):
        try:
            lib = _capi.load()
        except Exception:  # pragma: no cover - interpreter shutdown / library gone
            return
        for h in handles:
            try:
                lib.vp3d_plan_destroy(h)
            except Exception:  # pragma: no cover
                pass
        del handles[:]


def _default_precision():
    p = os.environ.get("VP3D_PRECISION", "fp16")
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
        self._train_precision = os.environ.get("VP3D_TRAIN_PRECISION", "bf16")
        self._plan = None
        self._plan_key = None
        self._plans = _PlanStore()
        self._packed = {}
        self._stats_epoch = 0      # bumped by every training forward (running stats changed)
        self._fwd_token = 0        # identifies the most recent training forward
        self._grad_reducer = None  # data_parallel.GradientReducer, set by its attach()
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
        self._register_params()

    def _register_params(self):
        ref = weakref.ref(self)
        for prm in self.parameters():
            _PARAM_OWNER[id(prm)] = ref

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
        """Eval-mode operand format.  'fp16' (default): IEEE fp16 operands and activations, fp32
        accumulate -- ~4e-4 of the fp32 reference at the tensor rate of bf16; 'bf16x3': every GEMM
        split-bf16 (fp32-faithful, ~1e-5); 'mixed': bf16 residual blocks on a hi+lo residual
        stream, split-bf16 expand / shrink (~1e-3); 'bf16': every GEMM plain bf16 (~3e-3).  Not in
        the reference."""
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self._precision = precision
        return self

    @property
    def precision(self):
        return self._precision

    def _reset_engine_state(self):
        """Forget every derived cache (plans, packed weights, workspace); the next forward rebuilds
        them.  The plan handles themselves are released when their store is collected."""
        self._plans = _PlanStore()
        self._packed = {}
        self._plan = None
        self._plan_key = None
        self._workspace = None

    def invalidate(self):
        """Force a re-pack of every parameter on the next forward.  Needed only after edits that
        bypass torch's version counters (``p.data.mul_()``, ``p.data.copy_()``, raw-pointer
        writes): ordinary in-place ops, ``optimizer.step`` and ``load_state_dict`` are detected
        automatically through ``tensor._version``.  Not in the reference."""
        self._packed = {}
        return self

    # copies / replicas never share engine state with the original (the handles point into one
    # device allocation each; sharing them made a collected copy free the original's plans)
    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        import copy as _copy
        skip = ("_plans", "_packed", "_plan", "_plan_key", "_workspace", "_grad_reducer")
        for k, v in self.__dict__.items():
            if k in skip:
                continue
            new.__dict__[k] = _copy.deepcopy(v, memo)
        new._reset_engine_state()
        new._grad_reducer = None
        new._register_params()
        return new

    def __copy__(self):
        cls = self.__class__
        new = cls.__new__(cls)
        new.__dict__.update(self.__dict__)
        # nn.Module.__copy__-style shallow copy of the registries
        for k in ("_parameters", "_buffers", "_modules"):
            new.__dict__[k] = self.__dict__[k].copy()
        new._reset_engine_state()
        return new

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("_plans", "_packed", "_plan", "_plan_key", "_workspace", "_grad_reducer"):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._reset_engine_state()
        self._grad_reducer = None
        self._register_params()

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica._reset_engine_state()
        return replica

    def _config(self, precision):
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
        cfg.precision = _PRECISIONS[precision]
        return cfg

    def _get_plan(self, device, precision=None):
        precision = precision or self._precision
        key = (device.index, precision)
        plans = self.__dict__.get("_plans")
        if plans is None:
            plans = self._plans = _PlanStore()
        if key not in plans:
            lib = _capi.load()
            handle = _capi.ctypes.c_void_p()
            cfg = self._config(precision)
            with torch.cuda.device(device):
                _capi.check(lib.vp3d_plan_create(_capi.ctypes.byref(cfg), _capi.ctypes.byref(handle)),
                            "vp3d_plan_create")
            plans.add(key, handle)
        self._plan = plans[key]
        self._plan_key = key
        return plans[key]

    def _param_tensors(self):
        """All fp32 tensors of the state_dict in the order of ``vp3d_weights``."""
        conv = [self.expand_conv.weight] + [c.weight for c in self.layers_conv] + \
               [self.shrink.weight]
        bn = []
        for m in [self.expand_bn] + list(self.layers_bn):
            bn += [m.weight, m.bias, m.running_mean, m.running_var]
        bn.append(self.shrink.bias)
        return conv, bn

    def _weights_struct(self):
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
        return w

    def _sync_weights(self, plan, stream, training=False):
        """Re-pack whatever changed since this plan last saw the parameters (optimizer.step,
        load_state_dict, in-place edits; the training kernels' running-stat updates are tracked
        through ``_stats_epoch`` because they bypass torch's version counters)."""
        conv, bn = self._param_tensors()
        versions = (tuple((t.data_ptr(), t._version) for t in conv),
                    tuple((t.data_ptr(), t._version) for t in bn) + (self._stats_epoch,))
        packed = self.__dict__.setdefault("_packed", {})
        key = (self._plan_key, training)
        seen = packed.get(key)
        what = 0
        if seen is None or seen[0] != versions[0]:
            what |= _capi.VP3D_PACK_CONV
            if training:
                what |= _capi.VP3D_PACK_CONV_T
        if not training and (seen is None or seen[1] != versions[1]):
            what |= _capi.VP3D_PACK_BN_EVAL
        if not what:
            return
        for t in conv + bn:
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("parameters must be contiguous float32 tensors")
        w = self._weights_struct()
        _capi.check(_capi.load().vp3d_set_weights(plan, _capi.ctypes.byref(w), what, stream),
                    "vp3d_set_weights")
        packed[key] = versions

    # -- hooks for optim.FusedAdam.attach (update + re-pack in one kernel) ---------------------------
    def _train_plan_ready(self, device):
        """True when the training plan on `device` already holds packed weights (i.e. a training
        forward has run): only then can the fused optimizer keep them current."""
        key = ((device.index, self._train_precision), True)
        return self.__dict__.get("_packed", {}).get(key) is not None

    def _mark_train_packs_current(self, device):
        """The fused optimizer step has just re-packed every conv weight of the training plan:
        record the parameters' new versions so that the next forward does not pack again."""
        conv, bn = self._param_tensors()
        versions = (tuple((t.data_ptr(), t._version) for t in conv),
                    tuple((t.data_ptr(), t._version) for t in bn) + (self._stats_epoch,))
        self._packed[((device.index, self._train_precision), True)] = versions

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
        params = self._learnable_tensors()
        return _TrainFunction.apply(self, x.contiguous(), *params)

    def _learnable_tensors(self):
        """Learnable tensors in the order of ``vp3d_grads``."""
        out = [self.expand_conv.weight, self.expand_bn.weight, self.expand_bn.bias]
        out += [c.weight for c in self.layers_conv]
        for m in self.layers_bn:
            out += [m.weight, m.bias]
        out += [self.shrink.weight, self.shrink.bias]
        return out

    def _learnable_names(self):
        out = ["expand_conv.weight", "expand_bn.weight", "expand_bn.bias"]
        out += [f"layers_conv.{i}.weight" for i in range(len(self.layers_conv))]
        for i in range(len(self.layers_bn)):
            out += [f"layers_bn.{i}.weight", f"layers_bn.{i}.bias"]
        out += ["shrink.weight", "shrink.bias"]
        return out

    def set_train_precision(self, precision):
        """'bf16' (default) or 'bf16x3' (fp32-faithful gradients) for the training kernels."""
        if precision not in ("bf16", "bf16x3"):
            raise ValueError("train precision must be 'bf16' or 'bf16x3'")
        self._train_precision = precision
        return self

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

    def forward_host_submit(self, x_host, out, slot):
        """Pipelined form of forward_host: enqueue copy-in -> kernels -> copy-out for one batch on
        `slot` (0 or 1) and return immediately; `forward_host_wait(slot)` completes it.  Alternating
        the slots overlaps the PCIe copy of the next batch with the kernels of the current one.
        `x_host` and `out` must be pinned CPU float32 tensors that stay alive until the wait."""
        lib = _capi.load()
        if self.training:
            raise RuntimeError("forward_host_submit is an eval-mode call")
        if x_host.is_cuda or x_host.dtype != torch.float32 or not x_host.is_contiguous():
            raise ValueError("forward_host_submit expects a contiguous float32 CPU tensor")
        assert x_host.dim() == 4 and x_host.shape[-2] == self.num_joints_in \
            and x_host.shape[-1] == self.in_features
        device = self.expand_conv.weight.device
        N, T = int(x_host.shape[0]), int(x_host.shape[1])
        with torch.cuda.device(device):
            plan = self._get_plan(device)
            stream = torch.cuda.current_stream(device)
            before = self._packed.get((self._plan_key, False))
            self._sync_weights(plan, stream.cuda_stream)
            if self._packed.get((self._plan_key, False)) is not before:
                stream.synchronize()  # freshly packed weights are read by the plan's own stream
            _capi.check(lib.vp3d_forward_eval_host_submit(plan, x_host.data_ptr(), out.data_ptr(), N,
                                                          T, int(slot)),
                        "vp3d_forward_eval_host_submit")
        return out

    def forward_host_wait(self, slot):
        _capi.check(_capi.load().vp3d_forward_eval_host_wait(self._plan, int(slot)),
                    "vp3d_forward_eval_host_wait")

    def last_launch_count(self):
        return 0 if self._plan is None else _capi.load().vp3d_last_launch_count(self._plan)


class _TrainFunction(torch.autograd.Function):
    """Training-mode forward/backward through the C ABI (vp3d_forward_train / vp3d_backward).

    The learnable tensors are passed as inputs so that autograd accumulates the returned
    gradients into ``.grad`` exactly as it does for the reference's nn modules."""

    @staticmethod
    def forward(ctx, module, x, *params):
        lib = _capi.load()
        device = x.device
        N, T = int(x.shape[0]), int(x.shape[1])
        momenta = [module.expand_bn.momentum] + [bn.momentum for bn in module.layers_bn]
        if any(m is None for m in momenta):
            raise NotImplementedError("BatchNorm momentum=None (cumulative average) is not supported")
        p_drop = float(module.drop.p)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # CPU generator: torch.manual_seed applies
        with torch.cuda.device(device):
            plan = module._get_plan(device, module._train_precision)
            stream = torch.cuda.current_stream(device).cuda_stream
            module._sync_weights(plan, stream, training=True)
            t_out = lib.vp3d_output_frames(plan, T)
            nbytes = lib.vp3d_train_workspace_bytes(plan, N, T)
            if t_out < 1 or nbytes == 0:
                raise ValueError(f"input of {T} frames is shorter than the receptive field "
                                 f"({module.receptive_field()})")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            y = torch.empty((N, t_out, module.num_joints_out, 3), dtype=torch.float32, device=device)
            w = module._weights_struct()
            mom = (_capi.ctypes.c_float * len(momenta))(*[float(m) for m in momenta])
            _capi.check(lib.vp3d_forward_train(plan, x.data_ptr(), y.data_ptr(), N, T,
                                               _capi.ctypes.byref(w), mom, p_drop, seed,
                                               ws.data_ptr(), ws.numel(), stream),
                        "vp3d_forward_train")
        # nn.BatchNorm1d bookkeeping that lives outside the kernels
        with torch.no_grad():
            module.expand_bn.num_batches_tracked += 1
            for bn in module.layers_bn:
                bn.num_batches_tracked += 1
        module._stats_epoch += 1
        module._fwd_token += 1
        ctx.module = module
        ctx.plan = plan
        ctx.ws = ws
        ctx.token = module._fwd_token
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.device = device
        return y

    @staticmethod
    def backward(ctx, dy):
        module = ctx.module
        if ctx.token != module._fwd_token:
            raise RuntimeError("only the most recent training forward of a module can be "
                               "back-propagated (one forward per backward, as in run.py)")
        lib = _capi.load()
        device = ctx.device
        dy = dy.contiguous().float()
        names = [n for n, _ in module.named_parameters()]
        order = module._learnable_names()
        reducer = getattr(module, "_grad_reducer", None)
        if reducer is not None and reducer.world > 1:
            # one flat buffer laid out in backward-completion order: each stage is a contiguous
            # slice that is all-reduced on a side stream while later stages are still computing
            _, spans, stage_spans, total = reducer.plan_layout(module)
            flat = torch.empty(total, dtype=torch.float32, device=device)
            by_name = {n: flat[spans[n][0]: spans[n][0] + spans[n][1]] for n in spans}
            grads = [by_name[n].view(s) for n, s in zip(order, ctx.shapes)]
        else:
            flat, stage_spans = None, None
            grads = [torch.empty(s, dtype=torch.float32, device=device) for s in ctx.shapes]
        nb2 = len(module.layers_conv)
        g = _capi.Grads()
        g.expand_conv_weight = grads[0].data_ptr()
        g.expand_bn[0] = grads[1].data_ptr()
        g.expand_bn[1] = grads[2].data_ptr()
        for i in range(nb2):
            g.layers_conv_weight[i] = grads[3 + i].data_ptr()
            g.layers_bn[i][0] = grads[3 + nb2 + 2 * i].data_ptr()
            g.layers_bn[i][1] = grads[3 + nb2 + 2 * i + 1].data_ptr()
        g.shrink_weight = grads[3 + 3 * nb2].data_ptr()
        g.shrink_bias = grads[3 + 3 * nb2 + 1].data_ptr()
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            if flat is None:
                _capi.check(lib.vp3d_backward(ctx.plan, dy.data_ptr(), _capi.ctypes.byref(g),
                                              ctx.ws.data_ptr(), ctx.ws.numel(), stream),
                            "vp3d_backward")
            else:
                errors = []

                def _stage(stage, _user):
                    try:
                        lo, hi = stage_spans[stage]
                        reducer.stage_ready(flat, lo, hi)
                    except Exception as e:  # never raise through the C frame
                        errors.append(e)

                cb = _capi.STAGE_FN(_stage)
                _capi.check(lib.vp3d_backward_staged(ctx.plan, dy.data_ptr(), _capi.ctypes.byref(g),
                                                     ctx.ws.data_ptr(), ctx.ws.numel(), stream,
                                                     _capi.ctypes.cast(cb, _capi.ctypes.c_void_p),
                                                     None), "vp3d_backward_staged")
                if errors:
                    raise errors[0]
                reducer.finish(flat)
        del names
        ctx.ws = None
        return (None, None) + tuple(grads)


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
