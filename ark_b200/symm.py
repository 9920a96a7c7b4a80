"""Symmetric (peer-mapped + NVSwitch-multicast) home of the flat buffers under data parallelism.

`csrc/dp_reduce.cu` (K11) reduces gradients in the switch, applies Adam to the 1/world slice a rank owns and multicasts
the new parameters to every rank.  For that the fp32 parameters, the fp32 gradients and the bf16 shadow of EVERY rank
must sit at the same offsets of one symmetric allocation that is also bound to a multicast object.  torch's
`torch.distributed._symmetric_memory` does the plumbing (cuMem allocation, handle exchange over the process group's
store, multicast binding); everything on the data path is this repository's kernel.
"""
from __future__ import annotations

import torch

FLAG_BYTES = 4096          # flag block of csrc/dp_reduce.cu (512 B used), in front of the buffers


def _rendezvous(buf, group):
    import torch.distributed._symmetric_memory as symm_mem
    try:
        hdl = symm_mem.rendezvous(buf, group)
    except Exception:       # older spellings want the group registered first / addressed by name
        symm_mem.enable_symm_mem_for_group(group.group_name)
        hdl = symm_mem.rendezvous(buf, group.group_name)
    mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
    if not mc:
        raise RuntimeError("the symmetric allocation has no NVSwitch multicast mapping")
    import torch.distributed as dist
    rank = dist.get_rank(group)
    bases = [int(p) for p in hdl.buffer_ptrs]
    off = int(getattr(hdl, "offset", 0) or 0)
    here = buf.data_ptr()
    if bases[rank] + off != here:
        if bases[rank] != here:
            raise RuntimeError(f"symmetric memory: cannot locate this rank's buffer ({bases[rank]:#x} + {off} vs {here:#x})")
        off = 0
    return hdl, mc + off, [b + off for b in bases]


class SymmBuf:
    """A symmetric multicast byte buffer (the gathered factor matrices of ops.dp_allgather_mc)."""

    def __init__(self, nbytes: int, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.buf = symm_mem.empty(int(nbytes), dtype=torch.uint8, device=device)
        self.hdl, self.mc, self.peers = _rendezvous(self.buf, group)
        self.hdl.barrier()
        torch.cuda.synchronize(device)


class SymmFlat:
    def __init__(self, numel: int, device, group):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.numel, self.group = int(numel), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        o_param, o_grad, o_shadow = FLAG_BYTES, FLAG_BYTES + 4 * numel, FLAG_BYTES + 8 * numel
        nbytes = FLAG_BYTES + 10 * numel
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.hdl, mc, self.peer_flags = _rendezvous(self.buf, group)
        off = 0
        self.mc_param, self.mc_grad, self.mc_shadow = mc + off + o_param, mc + off + o_grad, mc + off + o_shadow
        self.param = self.buf[o_param:o_grad].view(torch.float32)
        self.grad = self.buf[o_grad:o_shadow].view(torch.float32)
        self.shadow = self.buf[o_shadow:o_shadow + 2 * numel].view(torch.bfloat16)
        self.hdl.barrier()          # every rank has zeroed its flags before anyone's kernel can signal
        torch.cuda.synchronize(device)

    def buffers(self):
        return self.param, self.grad, self.shadow

    def owned(self, s, e):
        """Element range of [s, e) (multiples of 4) whose Adam state rank r maintains, for every r."""
        n4 = (e - s) // 4
        per = (n4 + self.world - 1) // self.world
        return [(s + 4 * min(n4, r * per), s + 4 * min(n4, (r + 1) * per)) for r in range(self.world)]
