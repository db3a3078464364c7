"""Context parallelism (Ulysses) for ONE image on several GPUs of a node - SURVEY.md §8f N2.

The reference (models/_modeling_parallel.py:41 `ContextParallelConfig`, hooks/context_parallel.py:129,220,
models/attention_dispatch.py:2504 `TemplatedUlyssesAttention`, transformers/transformer_flux.py:573 `_cp_plan`) shards the
token sequence over the ranks, and around every attention runs two all-to-all collectives: sequence-sharded q/k/v ->
head-sharded q/k/v over the full sequence, and the attention output back.

Here neither all-to-all exists as a separate step (no NCCL on the data path):
  * the QKV GEMM of rank r stores the columns of the heads rank d owns straight into d's joint QKV buffer over NVLink
    (b200_conv_gemm_args.y_peers: ONE launch whose output column blocks go to different peer mappings; the weight rows are
    permuted once so that a destination's [q | k | v] rows are one block, q / k normalised and rotated in the epilogue);
  * the attention kernel stores output row i straight into the buffer of the rank that owns row i
    (b200_attention_args.o_seg);
  * b200_peer_barrier (one tiny kernel, flags in peer memory) separates the phases.
The joint sequence is kept in RANK-MAJOR order ([rank 0: text slice, image slice | rank 1: ...]): attention is invariant
to a permutation of the keys once the rotary embedding has been applied with the tokens' true positions, and the query
order only permutes output rows - which are scattered back to their owners anyway.

`UlyssesPlan` is pure index arithmetic (tested on CPU with gloo, tests/test_context_parallel_cpu.py); `PeerGroup` owns the
peer-mapped buffers (GPU only).
"""
import ctypes as C
import dataclasses

import torch
import torch.distributed as dist

from . import _lib
from ._lib import B200Error


@dataclasses.dataclass
class ContextParallelConfig:
    """Same fields as the reference's ContextParallelConfig (models/_modeling_parallel.py:41); only Ulysses is built."""
    ring_degree: int = None
    ulysses_degree: int = None
    convert_to_fp32: bool = True
    rotate_method: str = "allgather"
    mesh: object = None
    ulysses_anything: bool = False
    ring_anything: bool = False

    def __post_init__(self):
        if self.ring_degree is None:
            self.ring_degree = 1
        if self.ulysses_degree is None:
            self.ulysses_degree = 1
        if self.ring_degree != 1 or self.ring_anything:
            raise NotImplementedError("ring attention is not built: use ulysses_degree (NVSwitch gives every GPU full all-to-all bandwidth)")
        if self.ulysses_anything:
            raise NotImplementedError("ulysses_anything (uneven splits) is not built: heads and both sequence lengths must divide by ulysses_degree")
        if self.mesh is not None:
            raise NotImplementedError("custom device meshes are not supported: the context-parallel group is the default process group")
        if self.ulysses_degree < 1 or self.ulysses_degree > 8:
            raise ValueError("ulysses_degree must be in [1, 8] (the GPUs of one NVSwitch node)")


class UlyssesPlan:
    """Who owns what: `world` ranks, a joint sequence of T text + S image tokens, `heads` heads of `head_dim`.

    rank r owns text rows [r*Tl, (r+1)*Tl), image rows [r*Sl, (r+1)*Sl) and heads [r*hl, (r+1)*hl).
    Rank-major joint order: rank r's block of Ll = Tl + Sl rows ([its text | its image]) starts at r*Ll."""

    def __init__(self, world, rank, text_len, image_len, heads, head_dim):
        if heads % world or text_len % world or image_len % world:
            raise ValueError(f"heads ({heads}), text length ({text_len}) and image length ({image_len}) must divide by the ulysses degree ({world})")
        self.world, self.rank = world, rank
        self.T, self.S, self.L = text_len, image_len, text_len + image_len
        self.Tl, self.Sl = text_len // world, image_len // world
        self.Ll = self.Tl + self.Sl
        if self.Tl % 8 or self.Sl % 8:
            raise ValueError("per-rank text and image slices must be multiples of 8 rows (16-byte aligned TMA rows)")
        self.heads, self.head_dim = heads, head_dim
        self.hl = heads // world
        self.D = heads * head_dim
        self.Dl = self.hl * head_dim

    # --- rows
    def text_rows(self, r=None):
        r = self.rank if r is None else r
        return r * self.Tl, (r + 1) * self.Tl

    def image_rows(self, r=None):
        r = self.rank if r is None else r
        return r * self.Sl, (r + 1) * self.Sl

    def block(self, r=None):
        """Rows of rank r's block in the rank-major joint order."""
        r = self.rank if r is None else r
        return r * self.Ll, (r + 1) * self.Ll

    def joint_to_rank_major(self):
        """index tensor g: rank-major row i holds the token at position g[i] of the reference's joint order [text | image]."""
        idx = []
        for r in range(self.world):
            t0, t1 = self.text_rows(r)
            s0, s1 = self.image_rows(r)
            idx.append(torch.arange(t0, t1))
            idx.append(self.T + torch.arange(s0, s1))
        return torch.cat(idx)

    # --- fused-QKV weight rows ([q | k | v], each D rows, head-major): destination-major order
    def qkv_row_permutation(self, order=None):
        """perm: row j of the permuted weight = row perm[j] of the fused [q | k | v] weight; block i (3*Dl rows) holds destination
        order[i]'s rows [q heads | k heads | v heads] (order = range(world) unless given; the model uses send_order(), so that at
        any moment the ranks' GEMMs store to different peers)."""
        idx = []
        for d in (range(self.world) if order is None else order):
            for part in range(3):
                idx.append(part * self.D + torch.arange(d * self.Dl, (d + 1) * self.Dl))
        return torch.cat(idx)

    def send_order(self):
        """Destinations starting with the next rank, so that at any moment the ranks write to different peers."""
        return [(self.rank + 1 + i) % self.world for i in range(self.world)]


class _RawDeviceMemory:
    """__cuda_array_interface__ holder: lets torch alias memory it did not allocate (a peer mapping)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerBuffer:
    """One buffer of `nbytes` on every rank of the group, each mapped into every process: `.local` is this rank's own
    (uint8 tensor), `.peer(r)` rank r's (over NVLink when r is another rank)."""

    def __init__(self, nbytes, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nbytes = (nbytes + 255) // 256 * 256
        dev = torch.cuda.current_device()
        _lib.init(dev)
        lib = _lib.lib()
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        # A failure on ONE rank (out of memory, no peer access) must not leave the others waiting in the next collective: every step
        # is followed by an exchange of its outcome, and all ranks raise together.
        err = None
        try:
            _lib.check(lib.b200_peer_alloc(self.nbytes, C.byref(ptr), handle), "b200_peer_alloc")
        except B200Error as e:
            err = str(e)
        self._own_ptr = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, (bytes(handle.raw), err), group=group)
        self._raise_together([e for _, e in handles], "allocate")
        handles = [h for h, _ in handles]
        self._ptrs, self._opened = [], []
        for r in range(self.world):
            if r == self.rank:
                self._ptrs.append(self._own_ptr)
                continue
            p = C.c_void_p()
            try:
                _lib.check(lib.b200_peer_open(handles[r], C.byref(p)), f"b200_peer_open (rank {r})")
            except B200Error as e:
                err = str(e)
                break
            self._ptrs.append(p.value)
            self._opened.append(p.value)
        outcomes = [None] * self.world
        dist.all_gather_object(outcomes, err, group=group)
        self._raise_together(outcomes, "map")
        self._holders = [_RawDeviceMemory(p, self.nbytes) for p in self._ptrs]
        self._tensors = [torch.as_tensor(h, device=torch.device("cuda", dev)) for h in self._holders]
        dist.barrier(group=group)  # every mapping exists before anyone writes through one

    def _raise_together(self, errors, what):
        bad = [(r, e) for r, e in enumerate(errors) if e]
        if bad:
            raise B200Error(f"context parallelism: could not {what} the peer buffers on rank(s) {[r for r, _ in bad]}: {bad[0][1]}")

    @property
    def local(self):
        return self._tensors[self.rank]

    def peer(self, r):
        return self._tensors[r]

    def ptr(self, r):
        return self._ptrs[r]

    def view(self, r, dtype, shape, byte_offset=0):
        n = 1
        for s in shape:
            n *= s
        nb = n * torch.empty((), dtype=dtype).element_size()
        if byte_offset % 256 or byte_offset + nb > self.nbytes:
            raise B200Error("PeerBuffer.view: out of range / misaligned")
        return self._tensors[r][byte_offset:byte_offset + nb].view(dtype).view(*shape)

    def close(self):
        if self._tensors is None:
            return
        torch.cuda.synchronize()
        self._tensors = self._holders = None
        lib = _lib.lib()
        dist.barrier(group=self.group)
        for p in self._opened:
            lib.b200_peer_close(C.c_void_p(p))
        dist.barrier(group=self.group)
        lib.b200_peer_free(C.c_void_p(self._own_ptr))


class PeerGroup:
    """The peer-mapped state of one context-parallel group: the flag words and epoch of b200_peer_barrier plus named
    data buffers carved out of one PeerBuffer per rank."""

    FLAG_BYTES = 256

    def __init__(self, data_bytes, group=None):
        self.buf = PeerBuffer(self.FLAG_BYTES + 256 + data_bytes, group)
        self.world, self.rank = self.buf.world, self.buf.rank
        self._flags = (C.c_void_p * self.world)(*[self.buf.ptr(r) for r in range(self.world)])
        self._epoch = torch.zeros(1, dtype=torch.int32, device="cuda")
        self._data_off = self.FLAG_BYTES + 256
        self._cursor = 0
        self.barriers = 0

    def carve(self, dtype, shape):
        """The same region of every rank's buffer as tensors: list indexed by rank."""
        n = 1
        for s in shape:
            n *= s
        nb = (n * torch.empty((), dtype=dtype).element_size() + 255) // 256 * 256
        off = self._data_off + self._cursor
        self._cursor += nb
        return [self.buf.view(r, dtype, shape, off) for r in range(self.world)]

    def barrier(self):
        """All ranks of the group: every write any of them issued on its stream before this call (to any rank's buffer) is
        visible to every kernel issued after it."""
        from . import ops
        _lib.check(_lib.lib().b200_peer_barrier(self._flags, C.c_void_p(self._epoch.data_ptr()), self.rank, self.world, ops._stream()),
                   "b200_peer_barrier")
        ops._count()
        self.barriers += 1

    def close(self):
        self.buf.close()
