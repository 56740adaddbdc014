"""Peer-memory mailboxes for the fused finalise + all-reduce kernel (include/cytof_b200.h,
`cytof_elbo_fwdbwd_allreduce`).  The reference has no communication; this is the one
exchange of a sharded step: the sum over ranks of the packed vector [grad block | ll | log_qy].

One process per GPU.  Every rank allocates a mailbox through the C ABI, the 64-byte CUDA-IPC
handles travel through `torch.distributed.all_gather_object` (any backend; plumbing only), and each
rank maps its peers' mailboxes.  After that a step needs no NCCL call: the finalise kernel stores
its outputs into every mailbox over NVLink / NVSwitch and sums the `world` slots in rank order."""
import ctypes as C

import torch
import torch.distributed as dist

from .. import _lib


class PeerMailboxes:
    def __init__(self, engine, group):
        self.lib = _lib.load()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.MAX_PEERS:
            raise ValueError('cytofpy_b200: at most %d ranks share mailboxes' % _lib.MAX_PEERS)
        self.device = engine.device
        d = engine._desc(engine.full_batch())
        nbytes = int(self.lib.cytof_peer_mailbox_bytes(C.byref(d), C.c_int32(self.world)))
        if nbytes == 0:
            raise RuntimeError('cytofpy_b200: descriptor rejected by cytof_peer_mailbox_bytes')
        self.ptrs = [None] * self.world
        # every step below that can fail is followed by an agreement over ALL ranks, so that a rank
        # that cannot map a peer makes every rank raise instead of leaving the others in a collective
        err, handle = None, C.create_string_buffer(64)
        try:
            with torch.cuda.device(self.device):
                own = C.c_void_p()
                _lib.check(self.lib.cytof_peer_alloc(C.c_size_t(nbytes), C.byref(own)))
                _lib.check(self.lib.cytof_ipc_get_handle(own, handle))
            self.ptrs[self.rank] = own.value
        except RuntimeError as e:
            err = str(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, None if err else bytes(handle.raw), group=group)
        if err is None and any(h is None for h in handles):
            err = 'a peer could not allocate / export its mailbox'
        if err is None:
            try:
                with torch.cuda.device(self.device):
                    for q in range(self.world):
                        if q == self.rank:
                            continue
                        ptr = C.c_void_p()
                        _lib.check(self.lib.cytof_ipc_open_handle(C.c_char_p(handles[q]), C.byref(ptr)))
                        self.ptrs[q] = ptr.value
            except RuntimeError as e:
                err = str(e)
        status = [None] * self.world
        dist.all_gather_object(status, err, group=group)    # also the barrier: every mailbox is mapped everywhere
        if any(st is not None for st in status):
            self.close()
            raise RuntimeError('cytofpy_b200: peer-memory mailboxes unavailable (%s); use allreduce="nccl"'
                               % next(st for st in status if st is not None))
        self.seq = 0
        self.desc = _lib.PeerDesc()
        self.desc.world, self.desc.rank = self.world, self.rank
        for q in range(self.world):
            self.desc.mailbox[q] = self.ptrs[q]

    def next(self):
        """descriptor of the next step (all ranks call this once per step, in lock step)"""
        self.seq += 1
        self.desc.seq = self.seq
        return self.desc

    def timeline(self):
        """ns stamps CTA 0 of the last exchange left in this rank's mailbox: (entry, own part stored, all flags seen,
        sum done) -- the split of the exchange cost that bench.py reports for N > 1"""
        out = (C.c_uint64 * 4)()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            _lib.check(self.lib.cytof_peer_debug(C.c_void_p(self.ptrs[self.rank]), out))
        return [int(v) for v in out]

    def close(self):
        if not self.ptrs:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for q, ptr in enumerate(self.ptrs):
                if ptr is None:
                    continue
                if q == self.rank:
                    self.lib.cytof_peer_free(C.c_void_p(ptr))
                else:
                    self.lib.cytof_ipc_close(C.c_void_p(ptr))
        self.ptrs = []
