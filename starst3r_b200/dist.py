"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on the B200 box, gloo in the CPU tests).

The reference is single-device; sharding is new capability:
  * MATCH: unordered image pairs are independent (sparse_ga.py:529 loop body) -> pair p goes to rank p mod G; the
    only communication is the exchange of the per-pair results afterwards (reconstruct.forward_mast3r does this
    under an initialised process group: reconstruct.SHARD_PAIRS).
  * RASTER training: camera views are sharded, the splat is replicated; per step ONE all-reduce (sum) of the
    per-Gaussian gradients (23 floats per Gaussian), then every rank applies the same Adam update.
  * ALIGN: does not shard (O(11 N) parameters, strictly sequential iterations): rank 0 runs the optimiser and
    broadcasts parameters and results (reconstruct._broadcast_alignment), so every rank continues from the same state.
"""
import os

import torch
import torch.distributed as dist

GRAD_KEYS = ("means", "quats", "scales", "opacities", "sh")


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items, rank=None, world_size=None):
    """Round-robin ownership: item i belongs to rank i mod G."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return list(range(rank, n_items, world_size))


def unordered_pairs(n_views):
    """The N(N-1)/2 pairs inference actually runs on (make_pairs 'complete' before symmetrisation)."""
    return [(i, j) for i in range(n_views) for j in range(i)]


def shard_pairs(n_views, rank=None, world_size=None):
    pairs = unordered_pairs(n_views)
    return [pairs[k] for k in shard_indices(len(pairs), rank, world_size)]


def allreduce_gradients(grads, group=None):
    """Sum the per-Gaussian gradient tensors over all ranks with a single collective (flatten -> all_reduce -> views)."""
    _, w = world()
    if w == 1:
        return grads
    flat = torch.cat([grads[k].reshape(-1) for k in GRAD_KEYS])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for k in GRAD_KEYS:
        n = grads[k].numel()
        grads[k].copy_(flat[off:off + n].view_as(grads[k]))
        off += n
    return grads


def gather_varlen(t, group=None):
    """All-gather of tensors whose first dimension differs per rank (correspondence lists): lengths first, then one
    padded all_gather.  Returns the list of per-rank tensors (on every rank)."""
    _, w = world()
    if w == 1:
        return [t]
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(out, pad, group=group)
    return [o[:s] for o, s in zip(out, sizes)]


def broadcast_tensors(tensors, src=0, group=None):
    _, w = world()
    if w > 1:
        for t in tensors:
            dist.broadcast(t, src, group=group)
    return tensors


class PeerGradExchange:
    """Gradient exchange of sharded-view training without a collective call: every rank's per-Gaussian gradients live
    in a symmetric-memory buffer that all ranks of the node map over NVLink (torch symmetric memory does the
    allocation, the handle exchange and the cross-device barrier: plumbing); the sum over ranks happens inside the
    fused Adam kernel (csrc/gs_adam.cu: st3r_adam_step_peers), which reads every peer's buffer directly.
    Two gradient buffers alternate between iterations, so ONE barrier per step orders both "all gradients of step k
    are written" and "nobody still reads the buffer step k+1 will overwrite"."""
    FLOATS = 23            # means 3 | quats 4 | scales 3 | opacity 1 | SH (4 coefficients x 3)
    # From this many ranks on the exchange runs as reduce-scatter + all-gather (st3r_grad_reduce_scatter: every rank sums
    # 1/G of the elements and stores them to all peers, 2 (G-1)/G L remote bytes) instead of every rank loading all
    # peers' gradients inside the Adam kernel ((G-1) L).  Same sums, bit for bit.
    SCATTER_FROM = 4

    def __init__(self, n_gauss, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.N = int(n_gauss)
        self.group = group if group is not None else dist.group.WORLD
        self.stride = (self.FLOATS * self.N + 63) // 64 * 64          # floats per parity, 256-byte aligned
        self.buf = symm_mem.empty(2 * self.stride, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.rank, self.world = self.hdl.rank, self.hdl.world_size
        self.step = 0
        N = self.N
        self.offsets = dict(means=0, quats=3 * N, scales=7 * N, opacities=10 * N, sh=11 * N)
        self.sets = []
        for parity in (0, 1):
            b = self.buf[parity * self.stride:(parity + 1) * self.stride]
            self.sets.append(dict(means=b[0:3 * N].view(N, 3), quats=b[3 * N:7 * N].view(N, 4),
                                  scales=b[7 * N:10 * N].view(N, 3), opacities=b[10 * N:11 * N],
                                  sh=b[11 * N:23 * N].view(N, 4, 3)))
        self.scatter = self.world >= int(os.environ.get("ST3R_SCATTER_FROM", self.SCATTER_FROM))
        # NVLS: both buffers also have a multicast address when the node's NVSwitch supports it; the reduce-scatter then
        # runs inside the switch (st3r_grad_reduce_multimem).  ST3R_NVLS=0 keeps the peer-load kernel.
        self.mc_grads = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.red = symm_mem.empty(self.stride, dtype=torch.float32, device=device)       # reduced gradients (all-gathered)
        self.red.zero_()
        self.red_hdl = symm_mem.rendezvous(self.red, self.group)
        r = self.red
        self.reduced = dict(means=r[0:3 * N].view(N, 3), quats=r[3 * N:7 * N].view(N, 4), scales=r[7 * N:10 * N].view(N, 3),
                            opacities=r[10 * N:11 * N], sh=r[11 * N:23 * N].view(N, 4, 3))
        self.mc_reduced = int(getattr(self.red_hdl, "multicast_ptr", 0) or 0)
        self.multimem = bool(self.mc_grads and self.mc_reduced) and os.environ.get("ST3R_NVLS", "1") != "0"
        self.hdl.barrier(channel=0)

    def reduce_scatter(self):
        """Sum this rank's slice over all ranks and store it to every rank's `reduced` buffer (call between two
        barriers: gradients complete before, reduced complete after)."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        w = self.world
        if self.multimem:
            off = (self.step & 1) * self.stride * 4
            with torch.cuda.device(self.buf.device):
                _lib.check(lib.st3r_grad_reduce_multimem(w, self.rank, ctypes.c_void_p(self.mc_grads + off),
                                                         ctypes.c_void_p(self.mc_reduced), ctypes.c_int64(self.stride),
                                                         _lib.stream_ptr()), "st3r_grad_reduce_multimem")
            return
        g = (ctypes.c_void_p * w)(*self.peer_bases())
        r = (ctypes.c_void_p * w)(*[int(p) for p in self.red_hdl.buffer_ptrs])
        with torch.cuda.device(self.buf.device):
            _lib.check(lib.st3r_grad_reduce_scatter(w, self.rank, g, r, ctypes.c_int64(self.stride), _lib.stream_ptr()),
                       "st3r_grad_reduce_scatter")

    def grads(self):
        """Gradient tensors the backward pass of the CURRENT step must write into."""
        return self.sets[self.step & 1]

    def peer_bases(self):
        """Device addresses of every rank's buffer of the current parity, as seen from this device."""
        off = (self.step & 1) * self.stride * 4
        return [int(p) + off for p in self.hdl.buffer_ptrs]

    def barrier(self):
        self.hdl.barrier(channel=0)

    def advance(self):
        self.step += 1
