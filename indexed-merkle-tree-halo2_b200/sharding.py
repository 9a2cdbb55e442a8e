"""Subtree sharding, one process per GPU over torch.distributed.

A depth-d tree over N = 2^k ranks is N independent depth-(d-k) subtrees (rank g owns leaves [g n/N, (g+1) n/N), a
consequence of level[l+1][i] = H(level[l][2i], level[l][2i+1]), /root/reference/src/utils.rs:43-47) plus a k-level cap.
The only exchange of the build is ONE all-gather of the N subtree roots (N x 32 bytes); every rank then builds the cap
redundantly. Queries shard by leaf owner (paths, preimages) or arbitrarily (folds / witness traces); low-leaf lookups
take one more all-gather of per-rank predecessor candidates.

With the `nccl` backend every exchange happens INSIDE the C library (csrc/imt_comm.cu: the engine gets its own NCCL
communicator via imt_comm_create — torch.distributed only carries the 128-byte NCCL id to the ranks — and the calls
below are one-line forwards to imt_sharded_*: the same entry points a Rust host binds). With `gloo` the same steps run
through the piecewise C calls with the exchanges staged through host memory by torch.distributed, which is what the CPU
tests drive. The engine argument only needs the Engine/Tree methods used here, so the tests can substitute an
oracle-backed double for the host-side logic.
"""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def attach_communicator(engine, group=None):
    """Gives `engine` its own NCCL communicator over the ranks of the torch.distributed group (imt_comm_create). torch only
    moves the 128-byte NCCL id from rank 0 to the others; every later collective is issued by the C library."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if engine.comm_info()[1] == world and world > 1:
        return
    uid = engine.comm_unique_id() if rank == 0 else bytes(128)
    dev = torch.device("cuda", engine.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    engine.comm_create(rank, world, bytes(t.cpu().numpy().tobytes()))


class ShardedTree:
    def __init__(self, engine, local_preimages=None, group=None, d_preimages=None, n_local=None):
        """local_preimages: this rank's (n_local, 3, 4) uint64 host array — or d_preimages: a CUDA tensor / device
        pointer with n_local leaves already resident. next_idx fields hold GLOBAL slot numbers."""
        import torch
        dist = _dist()
        self.engine, self.group = engine, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world & (self.world - 1):
            raise ValueError("the number of ranks must be a power of two")
        self.on_device = dist.get_backend(group) == "nccl"
        self._torch = torch
        self._cdev = torch.device("cuda", engine.device) if self.on_device else torch.device("cpu")
        if self.on_device:
            attach_communicator(engine, group)
        if d_preimages is not None:
            self.n_local = int(n_local)
            self.tree = (engine.sharded_build_from_leaves_dev(d_preimages, n_local) if self.on_device
                         else engine.build_from_leaves_dev(d_preimages, n_local))
        else:
            self.n_local = int(np.asarray(local_preimages).reshape(-1, 3, 4).shape[0])
            self.tree = engine.sharded_build_from_leaves(local_preimages) if self.on_device else engine.build_from_leaves(local_preimages)
        if not self.on_device:
            self.tree.set_shard(self.rank, self.world)
            self.exchange_roots()

    # ---- build
    def exchange_roots(self):
        """all-gather of the N subtree roots + the replicated cap levels; call again after every rebuild"""
        if self.on_device:
            self.tree.exchange_roots()          # imt_tree_exchange_roots: ncclAllGather inside the library
        else:
            roots = self._all_gather(self.tree.subtree_root())
            self.tree.attach_cap(self.rank, self.world, roots)

    def rebuild(self, local_preimages=None, d_preimages=None):
        if self.on_device:
            if d_preimages is not None:
                self.tree.sharded_rebuild_from_leaves_dev(d_preimages)
            else:
                self.tree.sharded_rebuild_from_leaves(local_preimages)
            return
        if d_preimages is not None:
            self.tree.rebuild_from_leaves_dev(d_preimages)
        else:
            self.tree.rebuild_from_leaves(local_preimages)
        self.exchange_roots()

    @property
    def depth(self):
        return self.tree.depth

    @property
    def num_leaves(self):
        return self.n_local * self.world

    def root(self):
        return self.tree.root()

    # ---- collectives on numpy arrays (uint64 / uint8), staged through the backend's device
    def _all_gather(self, a):
        torch, dist = self._torch, _dist()
        a = np.ascontiguousarray(a)
        t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(self._cdev)
        t = t.contiguous().unsqueeze(0)  # gloo splits the output along dim 0 and wants each piece shaped like the input
        out = torch.empty((self.world,) + tuple(t.shape[1:]), dtype=t.dtype, device=self._cdev)
        dist.all_gather_into_tensor(out, t, group=self.group)
        r = out.cpu().numpy()
        return r.view(np.uint64) if a.dtype == np.uint64 else r

    def owner(self, indices):
        return (np.asarray(indices, dtype=np.uint64) // np.uint64(self.n_local)).astype(np.int64)

    def _served_by_owner(self, indices, serve, shapes):
        """Every rank holds the same `indices`; rank r computes serve(indices it owns) -> tuple of arrays, the pieces are
        all-gathered (padded to the largest piece) and reassembled in query order on every rank."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
        if idx.size and int(idx.max()) >= self.num_leaves:
            raise IndexError("index out of bounds")
        own = self.owner(idx)
        counts = np.bincount(own, minlength=self.world)
        pad = int(counts.max()) if idx.size else 0
        mine = np.nonzero(own == self.rank)[0]
        pieces = serve(idx[mine]) if mine.size else None
        outs = []
        for k, (inner, dtype) in enumerate(shapes):
            buf = np.zeros((pad,) + inner, dtype)
            if mine.size:
                buf[: mine.size] = pieces[k]
            allp = self._all_gather(buf) if pad else np.zeros((self.world, 0) + inner, dtype)
            full = np.empty((idx.size,) + inner, dtype)
            for r in range(self.world):
                full[own == r] = allp[r, : counts[r]]
            outs.append(full)
        return outs

    # ---- queries
    def get_proofs(self, indices):
        """get_proof (utils.rs:63-85) for GLOBAL leaf indices: the bottom d-k siblings come from the owner's subtree,
        the top k from the replicated cap. Returns (siblings [q, d, 4], helpers [q, d]) on every rank."""
        if self.on_device:
            return self._native(lambda: self.tree.sharded_get_proofs(indices))
        d = self.depth
        return tuple(self._served_by_owner(indices, lambda ix: self.tree.get_proofs(ix), [((d, 4), np.uint64), ((d,), np.uint8)]))

    def leaves(self, indices):
        if self.on_device:
            return self._native(lambda: self.tree.sharded_leaves(indices))
        return tuple(self._served_by_owner(indices, lambda ix: self.tree.leaves(ix), [((3, 4), np.uint64), ((), np.uint8)]))

    @staticmethod
    def _native(call):
        """the C library reports an out-of-range index as ImtError(INDEX_OOB); the gloo path raises IndexError"""
        from .engine import ImtError
        from . import _ffi
        try:
            return call()
        except ImtError as e:
            if e.status == _ffi.ERR_INDEX_OOB:
                raise IndexError("index out of bounds") from e
            raise

    def low_leaf_lookup(self, values):
        """update_idx_leaf's scan (IMT:632-660) over the sharded tree: per-rank predecessor candidates from each rank's
        sorted index, one all-gather, then the same decision as the single-GPU lookup. Replicated result."""
        v = np.ascontiguousarray(values, dtype=np.uint64).reshape(-1, 4)
        if self.on_device:
            return self.tree.sharded_low_leaf_lookup(v)      # candidates + ncclAllGather + merge inside the library
        keys, slots, flags = self.tree.low_leaf_candidates(v)
        occ = self._all_gather(np.array([self.tree.occupied, 1 if (self.rank == 0 and self.tree.head_next_zero) else 0], np.uint64))
        gk, gs, gf = self._all_gather(keys), self._all_gather(slots), self._all_gather(flags)
        return self.engine.low_leaf_merge(v, gk, gs, gf, int(occ[:, 0].sum()), self.num_leaves, bool(occ[0, 1]))

    def non_inclusion_paths(self, values):
        """witnesses of verify_non_inclusion (IMT:127-137) for every value, replicated on every rank"""
        if self.on_device:
            o = self.tree.sharded_non_inclusion_paths(values)
            o["matched"] = o["matched"].astype(bool)
            return o
        low, matched = self.low_leaf_lookup(values)
        leaves, largest = self.leaves(low)
        sib, hel = self.get_proofs(low)
        return dict(low_idx=low, matched=matched, low_leaves=leaves, siblings=sib, helpers=hel, is_largest=largest)

    def insert_batch(self, new_vals, chunk=4096):
        """imt_insert_batch over the sharded tree (IMT:710-741 with O(depth) hashes per insert): every rank passes the same
        values; the witnesses (same dict as Tree.insert_batch) come back replicated. Per chunk: neighbours from every
        rank's index -> all-gather -> replicated plan -> every rank applies its own writes to its subtree -> all-gather
        of the subtree-root versions and local paths -> every rank applies all writes to the replicated cap."""
        v = np.ascontiguousarray(new_vals, dtype=np.uint64).reshape(-1, 4)
        if self.on_device:
            from .engine import ImtError
            from . import _ffi
            try:
                return self.tree.sharded_insert_batch(v)     # the same rounds, exchanged by NCCL inside the library
            except ImtError as e:
                if e.status == _ffi.ERR_TREE_FULL:
                    raise ValueError("not enough empty slots") from e
                raise
        b_total, d = v.shape[0], self.depth
        d_local = self.n_local.bit_length() - 1
        d_cap = d - d_local
        keys = ("old_roots", "low_idx", "low_leaves", "low_siblings", "low_helpers", "new_roots", "new_leaves", "new_siblings", "new_helpers",
                "is_largest")
        parts = {k: [] for k in keys}
        occupied = int(self._all_gather(np.array([self.tree.occupied], np.uint64)).sum())
        for off in range(0, b_total, chunk):
            vals = v[off:off + chunk]
            b = vals.shape[0]
            if occupied + b > self.num_leaves:
                raise ValueError("not enough empty slots")
            root_before = self.root()
            gathered = [self._all_gather(a) for a in self.tree.shard_insert_neighbors(vals)]
            x, upd, low_old, largest = self.engine.shard_insert_plan(vals, occupied, *gathered)
            sub_roots, sib_local = self.tree.shard_insert_apply(x, upd, d_local)
            own = self.owner(x)
            t = np.arange(2 * b)
            sub_roots = self._all_gather(sub_roots)[own, t]
            sib_local = self._all_gather(sib_local)[own, t] if d_local else np.zeros((2 * b, 0, 4), np.uint64)
            roots, sib_cap = self.tree.shard_insert_cap(x, sub_roots, d_cap)
            sib = np.concatenate([sib_local, sib_cap], axis=1)
            helpers = (((x[:, None] >> np.arange(d, dtype=np.uint64)[None, :]) & np.uint64(1)) == 0).astype(np.uint8)
            parts["old_roots"].append(np.concatenate([root_before[None, :], roots[1:-1:2]]))
            parts["new_roots"].append(roots[1::2])
            parts["low_idx"].append(x[0::2])
            parts["low_leaves"].append(low_old)
            parts["new_leaves"].append(upd[1::2])
            parts["low_siblings"].append(sib[0::2])
            parts["new_siblings"].append(sib[1::2])
            parts["low_helpers"].append(helpers[0::2])
            parts["new_helpers"].append(helpers[1::2])
            parts["is_largest"].append(largest)
            occupied += b
        return {k: np.concatenate(p) if p else None for k, p in parts.items()}

    def query_slice(self, q):
        """by-query sharding of pure-compute batches (path folds, witness traces): this rank's slice of q queries"""
        per = -(-q // self.world)
        return slice(min(q, self.rank * per), min(q, (self.rank + 1) * per))

    def trace_proofs(self, indices):
        """verify_merkle_proof witness traces (IMT:65-96) for GLOBAL leaf indices, sharded by leaf owner with NO exchange:
        the owner reads every operand from its stored subtree levels / the replicated cap and runs the q x depth traced
        hashes independently (imt_tree_trace_proofs). Returns (mine, states): `mine` = positions into `indices` of the
        queries this rank owns, states[len(mine), depth, states per hash, t, 4] their traces — they stay on this rank."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
        if idx.size and int(idx.max()) >= self.num_leaves:
            raise IndexError("index out of bounds")
        mine = np.nonzero(self.owner(idx) == self.rank)[0]
        return mine, self.tree.trace_proofs(idx[mine])

    def trace_merkle_proofs(self, leaves, indices, siblings, want_states=True):
        """compute_merkle_root traces (IMT:78-96) of this rank's query slice: (slice, roots, states). The traces stay
        on the rank that produced them (19.9 GB for 2^16 depth-24 paths is not gathered)."""
        lv = np.asarray(leaves, dtype=np.uint64).reshape(-1, 4)
        sl = self.query_slice(lv.shape[0])
        if sl.start >= sl.stop:
            return sl, np.zeros((0, 4), np.uint64), None
        idx = np.asarray(indices, dtype=np.uint64).reshape(-1)[sl]
        sib = np.asarray(siblings, dtype=np.uint64).reshape(lv.shape[0], -1, 4)[sl]
        roots, states = self.engine.trace_merkle_proofs(lv[sl], idx, sib, want_states=want_states)
        return sl, roots, states
