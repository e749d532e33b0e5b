"""Cross-rank BatchNorm statistics (the reference converts every BatchNorm to SyncBatchNorm when more than one GPU is
used: main_utils.py:335-338, `nn.SyncBatchNorm.convert_sync_batchnorm(model)`).

The kernels of this package never call a BatchNorm module's forward: they take per-channel sums over the local rows
([sum z, sum z^2] forward; [sum dy, sum dy*zhat] backward), finalise them into scale / shift, and apply those in GEMM
prologues / epilogues.  Synchronising a layer therefore means summing those tiny vectors over the ranks between the
"stats" and the "finalise / apply" kernels — 2C numbers per layer and direction — and multiplying the row count by the
world size.  `reducer_of(bn)` decides whether a layer is synchronised:

  * the module is an `nn.SyncBatchNorm` (what the reference's unchanged conversion call produces on a model built from
    this package's modules), in training mode, and torch.distributed is initialised with world size > 1; or
  * `convert_sync_batchnorm(model)` of this module tagged it (same effect, keeps the module class).

The reduction itself is either torch.distributed's all-reduce on the current stream (NCCL on the GPUs, gloo in the CPU
tests) or — `enable_peer_reduce` — this package's own one-launch NVLink peer-memory exchange (csrc/peer_reduce.cu):
every rank stores its vector into every peer's slot, raises a flag there, waits for the peers' flags and sums the
slots in rank order, so all ranks obtain bit-identical statistics without a library collective on the critical path.

Parameter gradients of the BatchNorm affine terms stay LOCAL sums (DistributedDataParallel averages them with the rest
of the gradients), exactly as torch's SyncBatchNorm does.
"""
import torch
import torch.distributed as dist

_reducers = {}


class Reducer:
    """Sum small device vectors over the ranks of `group`, in place, on the current stream."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.peer = None            # PeerReduce once enabled (CUDA only)
        self._checked = set()

    def total_count(self, local_count):
        """Rows the statistics are taken over, all ranks together.  The kernels take the count as a host scalar, so
        every rank must contribute the same number of rows (DistributedSampler guarantees it); verified once per
        distinct local count with a host all-gather (outside CUDA-graph capture: warm-up steps come first)."""
        key = float(local_count)
        if key not in self._checked:
            if not (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()):
                counts = [None] * self.world
                dist.all_gather_object(counts, key, group=self.group)
                if any(c != key for c in counts):
                    raise RuntimeError(f"eda_b200.syncbn: ranks hold different row counts {counts}; synchronised "
                                       "BatchNorm here needs equal per-rank batches")
            self._checked.add(key)
        return float(local_count) * self.world

    def all_reduce_sum_(self, t):
        if self.peer is not None and t.is_cuda:
            self.peer.all_reduce_sum_(t)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def default_reducer(group=None):
    key = id(group) if group is not None else None
    r = _reducers.get(key)
    if r is None:
        r = _reducers[key] = Reducer(group)
    return r


def reducer_of(bn):
    """The Reducer that synchronises `bn`'s batch statistics, or None (single process, eval mode, plain BatchNorm)."""
    if bn is None or not bn.training or not _active():
        return None
    if isinstance(bn, torch.nn.SyncBatchNorm):
        return default_reducer(getattr(bn, "process_group", None))
    tag = bn.__dict__.get("_eda_sync_group", False)
    if tag is False:
        return None
    return default_reducer(tag)


def convert_sync_batchnorm(module, process_group=None):
    """Marks every BatchNorm1d / BatchNorm2d of `module` as synchronised over `process_group` (None = the default
    group).  Equivalent in effect to nn.SyncBatchNorm.convert_sync_batchnorm for the layers this package's kernels
    evaluate (which also works, see the module docstring) but keeps the module classes — and therefore isinstance checks
    and state-dict layout — untouched.  Returns `module`."""
    for m in module.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.__dict__["_eda_sync_group"] = process_group
    return module


def enable_peer_reduce(device, group=None, max_floats=4096):
    """Switches the default reducer of `group` to the NVLink peer-memory exchange (one kernel per reduction).  Needs
    torch's symmetric-memory rendezvous between the ranks of one node; returns False (and leaves NCCL in charge) when
    that is unavailable."""
    from . import peer

    red = default_reducer(group)
    try:
        red.peer = peer.PeerReduce(device, group, max_floats)
        return True
    except Exception as e:  # noqa: BLE001
        red.peer = None
        red.peer_error = repr(e)
        return False
