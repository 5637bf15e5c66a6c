"""Data-parallel training of the temporal models: one process per GPU, batches sharded by rank,
one gradient all-reduce per step over NCCL (NVLink 5 / NVSwitch) — SURVEY.md §8e.

The reference is single-GPU (no torch.distributed anywhere), so this module has no reference
counterpart; it adds exactly one collective.  BatchNorm statistics stay per-GPU (north_star:
"allreduce on gradients only"); `broadcast_buffers` aligns the running statistics before a
checkpoint / evaluation, as DistributedDataParallel does.

Overlap: the C backward (`vp3d_backward_staged`) reports, stage by stage, when the kernels producing
a group of gradients have been enqueued (shrink first, expand last).  All gradients of a step live
in one flat fp32 buffer laid out in that completion order, so each stage is a contiguous slice whose
all-reduce is launched on a side stream behind an event while the remaining backward GEMMs run.
"""
import torch
import torch.distributed as dist


def stage_order(module):
    """Parameter names grouped by backward completion stage (see vp3d_backward_staged)."""
    nb = len(module.layers_conv) // 2
    stages = [["shrink.weight", "shrink.bias"]]
    for i in range(nb, 0, -1):
        c1, c2 = 2 * (i - 1), 2 * (i - 1) + 1
        stages.append([f"layers_conv.{c2}.weight", f"layers_bn.{c2}.weight", f"layers_bn.{c2}.bias",
                       f"layers_conv.{c1}.weight", f"layers_bn.{c1}.weight", f"layers_bn.{c1}.bias"])
    stages.append(["expand_conv.weight", "expand_bn.weight", "expand_bn.bias"])
    return stages


class GradientReducer:
    """Averages gradients across the ranks of `process_group`.

    attach(module) makes the module's training backward write its gradients into a flat buffer and
    all-reduce it stage by stage (overlapped on CUDA); `reduce_flat` is the device-agnostic core and
    is what the CPU (gloo) tests exercise."""

    def __init__(self, process_group=None, overlap=True, compress=None, reserve_sms=0):
        """compress: None (fp32 all-reduce) or "bf16" (each slice is rounded to bf16 for the wire
        and widened back: half the NVLink bytes; the rounding error, 2^-9 relative per rank
        contribution, is below the bf16 training noise floor).
        reserve_sms: with overlap, cap the persistent GEMM grids at (SM count - reserve_sms) so
        that the NCCL kernels of the side stream always find free SMs (pair it with
        NCCL_MAX_CTAS <= reserve_sms in the environment)."""
        if compress not in (None, "bf16"):
            raise ValueError("compress must be None or 'bf16'")
        self.reserve_sms = int(reserve_sms)
        self.group = process_group
        self.overlap = overlap
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.comm_stream = None
        self.launched = 0
        self.step_scale = 1.0
        self.compress = compress

    # ---------------------------------------------------------------- layout
    def plan_layout(self, module):
        """-> (names in flat order, {name: (offset, numel)}, [(lo, hi)] per stage); offsets are
        padded to 4 elements so that every slice stays 16-byte aligned."""
        params = dict(module.named_parameters())
        names, spans, stage_spans = [], {}, []
        off = 0
        for group in stage_order(module):
            lo = off
            for n in group:
                numel = params[n].numel()
                spans[n] = (off, numel)
                names.append(n)
                off += (numel + 3) // 4 * 4
            stage_spans.append((lo, off))
        return names, spans, stage_spans, off

    def attach(self, module):
        """Route `module`'s training backward through this reducer; returns the reducer.

        With more than one rank on CUDA devices the kernels' programmatic dependent launch is
        switched off for the process: kernels that hand every SM over without a gap leave the NCCL
        kernels nowhere to run (measured on 2 x B200: 2.76 ms / step with it off; with it on the
        step is as fast only in one of the two regimes -- launch queue full or a host sync every
        step, as run.py's `loss.item()` does -- and collapses to tens of ms in the other; since
        the small kernels between the GEMMs also chain programmatically this holds without
        overlap too)."""
        object.__setattr__(module, "_grad_reducer", self)
        if self.world > 1 and torch.cuda.is_available():
            from . import _capi
            _capi.check(_capi.load().vp3d_set_pdl(0), "vp3d_set_pdl")
            if self.overlap and self.reserve_sms > 0:
                sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
                limit = (sms - self.reserve_sms) & ~1   # even: the GEMMs run on CTA pairs
                if limit >= 2:
                    _capi.check(_capi.load().vp3d_set_sm_limit(limit), "vp3d_set_sm_limit")
        return self

    def set_step_rows(self, local_rows, global_rows):
        """Weight this rank's gradient by its share of the step's batch rows.  With equal shards
        (the usual case) the weight is 1 and the all-reduce is a plain mean; with a ragged last
        batch the mean of per-rank means would over-weight the short shards, so each rank's
        gradient is scaled by n_local * world / n_global before the average."""
        if global_rows <= 0:
            raise ValueError("global_rows must be positive")
        self.step_scale = float(local_rows) * self.world / float(global_rows)

    # ---------------------------------------------------------------- collective
    def reduce_flat(self, flat):
        """In-place (weighted) average of a flat gradient slice over the group."""
        if self.world == 1:
            return flat
        if self.step_scale != 1.0:
            flat.mul_(self.step_scale)
        wire = flat.to(torch.bfloat16) if self.compress == "bf16" else flat
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            dist.all_reduce(wire, op=dist.ReduceOp.AVG, group=self.group)
        else:  # gloo (CPU tests): no AVG
            dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group)
            wire.div_(self.world)
        if wire is not flat:
            flat.copy_(wire)
        self.launched += 1
        return flat

    def stage_ready(self, flat, lo, hi):
        """Called (from the C backward's stage callback) once the kernels writing flat[lo:hi] are
        enqueued on the current stream."""
        if hi <= lo:
            return
        piece = flat[lo:hi]
        if not flat.is_cuda or not self.overlap:
            self.reduce_flat(piece)
            return
        cur = torch.cuda.current_stream(flat.device)
        if self.comm_stream is None:
            # (high priority: the collective's kernels must not queue behind the compute grids)
            self.comm_stream = torch.cuda.Stream(device=flat.device, priority=-1)
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            self.reduce_flat(piece)
        flat.record_stream(self.comm_stream)

    def finish(self, flat):
        """Make the current stream wait for every outstanding all-reduce."""
        if flat.is_cuda and self.comm_stream is not None:
            torch.cuda.current_stream(flat.device).wait_stream(self.comm_stream)


def broadcast_buffers(module, src=0, process_group=None):
    """Copy rank `src`'s BatchNorm running statistics to every rank (checkpoint / eval alignment)."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    for b in module.buffers():
        dist.broadcast(b, src=src, group=process_group)
    if hasattr(module, "_stats_epoch"):
        module._stats_epoch += 1


def shard_batch(batch_index, rank, world):
    """Weak scaling as SURVEY.md §8e prefers: rank r takes batches b with b % world == r, each of
    the full per-GPU batch size (keeps the BatchNorm population per GPU equal to the reference's)."""
    return batch_index % world == rank
