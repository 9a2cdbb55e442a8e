"""Batched, numpy-facing wrapper of the C-ABI: Engine (one context = one GPU) and Tree (device-resident levels).

Field elements are rows of 4 little-endian uint64 words (32 bytes) in the engine's format (canonical by default).
"""
import ctypes
import weakref

import numpy as np

from . import _ffi

P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
STATES_PER_HASH = 132


class ImtError(Exception):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def _fe_array(a, inner):
    """contiguous uint64 array whose trailing dims are `inner` + (4,)"""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a.reshape((-1,) + tuple(inner) + (4,))


def fe_from_int(x):
    x = int(x)
    if not 0 <= x < (1 << 256):
        raise ValueError("field element out of range")
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def fe_to_int(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1)
    return sum(int(a[i]) << (64 * i) for i in range(4))


def fes_from_ints(xs):
    return np.stack([fe_from_int(x) for x in xs]) if len(xs) else np.zeros((0, 4), np.uint64)


def fes_to_ints(a):
    return [fe_to_int(r) for r in np.asarray(a, dtype=np.uint64).reshape(-1, 4)]


def _dev_ptr(t):
    """device pointer of a torch CUDA tensor (or a raw int)"""
    if isinstance(t, int):
        return ctypes.c_void_p(t)
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    return ctypes.c_void_p(t.data_ptr())


class Engine:
    """One imt_ctx: Poseidon::<Fr,T,RATE>::new(r_f, r_p) + a GPU. Not thread-safe (mirrors the reference's &mut hasher).

    The default is the reference's instance <3, 2>(8, 57) on the tuned kernels (imt_ctx_create). Any other instance
    (t in 2..5, rate = t - 1) — or the same one with generic=True — runs every hash on the any-width kernels
    (imt_ctx_create_spec): utils.rs:6, 19 and indexed_merkle_tree.rs:65, 127, 231 are generic over T and RATE."""

    def __init__(self, device=0, fmt="canonical", t=3, rate=2, r_f=8, r_p=57, generic=False):
        self._lib = _ffi.load()
        self.fmt = {"canonical": _ffi.FE_CANONICAL, "montgomery": _ffi.FE_MONTGOMERY}[fmt]
        h = ctypes.c_void_p()
        self.t, self.rate, self.r_f, self.r_p = int(t), int(rate), int(r_f), int(r_p)
        self.generic = bool(generic) or (self.t, self.rate, self.r_f, self.r_p) != (3, 2, 8, 57)
        if self.generic:
            st = self._lib.imt_ctx_create_spec(int(device), self.fmt, self.t, self.rate, self.r_f, self.r_p, ctypes.byref(h))
            if st == _ffi.ERR_INVALID_ARG:
                raise ImtError(st, f"unsupported Poseidon instance T={t} RATE={rate} R_F={r_f} R_P={r_p}: need 2 <= T <= 5, "
                                   "RATE = T - 1, R_F even, R_F + R_P <= 256")
        else:
            st = self._lib.imt_ctx_create(int(device), self.fmt, ctypes.byref(h))
        if st != _ffi.OK:
            raise ImtError(st, f"imt_ctx_create(device={device}) failed with status {st}: a CUDA device is required, "
                               "there is no CPU fallback")
        self._h = h
        self._trees = weakref.WeakSet()   # a tree holds a pointer to its context: close() destroys the trees first
        self.device = int(device)
        self.states_per_perm = 1 + self.r_f + self.r_p

    def states_per_hash(self, arity):
        """traced states (of t FE each) of one hash of `arity` inputs: (arity // rate + 1) permutations x (1 + r_f + r_p)"""
        return (arity // self.rate + 1) * self.states_per_perm

    def close(self):
        if getattr(self, "_h", None):
            for t in list(getattr(self, "_trees", ())):
                t.close()
            if not getattr(self, "_borrowed", False):
                self._lib.imt_comm_destroy(self._h)
                self._lib.imt_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != _ffi.OK:
            msg = self._lib.imt_last_error(self._h).decode() or self._lib.imt_status_string(st).decode()
            raise ImtError(st, msg)

    @property
    def launches(self):
        return int(self._lib.imt_ctx_launch_count(self._h))

    def set_stream(self, cuda_stream_ptr):
        """launch on a caller-owned cudaStream_t (int handle; 0 = the legacy default stream, torch's default)"""
        self._check(self._lib.imt_ctx_set_stream(self._h, ctypes.c_void_p(int(cuda_stream_ptr or 0))))

    def reset_stream(self):
        """back to the engine's own non-blocking stream"""
        self._check(self._lib.imt_ctx_reset_stream(self._h))

    def trim(self):
        """give the cached scratch memory of earlier calls back to the driver"""
        self._check(self._lib.imt_ctx_trim(self._h))

    def enable_timing(self, on=True):
        self._check(self._lib.imt_ctx_enable_timing(self._h, 1 if on else 0))

    def reset_timing(self):
        self._check(self._lib.imt_ctx_reset_timing(self._h))

    def kernel_time(self, arity):
        """(total device ms, launches, hashes) of the hash kernels of this arity since the last reset"""
        ms, ln, hs = ctypes.c_double(), ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self._lib.imt_ctx_kernel_time(self._h, arity, ctypes.byref(ms), ctypes.byref(ln), ctypes.byref(hs)))
        return ms.value, ln.value, hs.value

    # ---- batched hashing
    def hash2(self, pairs):
        a = _fe_array(pairs, (2,))
        out = np.empty((a.shape[0], 4), np.uint64)
        self._check(self._lib.imt_poseidon_hash2(self._h, _ptr(a), a.shape[0], _ptr(out)))
        return out

    def hash3(self, triples):
        a = _fe_array(triples, (3,))
        out = np.empty((a.shape[0], 4), np.uint64)
        self._check(self._lib.imt_poseidon_hash3(self._h, _ptr(a), a.shape[0], _ptr(out)))
        return out

    def hash(self, inputs, arity, n=None):
        """update(&inputs[i]) + squeeze_and_reset() for any input length (arity >= 0); inputs: (n, arity, 4).
        With arity 0 (squeeze of an empty sponge) pass the number of hashes as n."""
        a = np.ascontiguousarray(inputs, dtype=np.uint64).reshape(-1 if arity else (n or 1), arity, 4)
        out = np.empty((a.shape[0], 4), np.uint64)
        self._check(self._lib.imt_poseidon_hash(self._h, _ptr(a) if arity else None, arity, a.shape[0], _ptr(out)))
        return out

    def convert(self, a, to_montgomery):
        """dense FE array canonical <-> Montgomery (imt_fe_convert), independent of the engine's own format"""
        a = np.ascontiguousarray(a, dtype=np.uint64)
        out = np.empty_like(a)
        frm, to = (_ffi.FE_CANONICAL, _ffi.FE_MONTGOMERY) if to_montgomery else (_ffi.FE_MONTGOMERY, _ffi.FE_CANONICAL)
        self._check(self._lib.imt_fe_convert(self._h, _ptr(a), a.size // 4, frm, to, _ptr(out)))
        return out

    def convert_dev(self, d_in, n, d_out, to_montgomery):
        frm, to = (_ffi.FE_CANONICAL, _ffi.FE_MONTGOMERY) if to_montgomery else (_ffi.FE_MONTGOMERY, _ffi.FE_CANONICAL)
        self._check(self._lib.imt_fe_convert_dev(self._h, _dev_ptr(d_in), n, frm, to, _dev_ptr(d_out)))

    def hash_dev(self, d_in, arity, n, d_out):
        self._check(self._lib.imt_poseidon_hash_dev(self._h, _dev_ptr(d_in), arity, n, _dev_ptr(d_out)))

    def permute(self, states):
        """the bare permutation on (n, t, 4) states"""
        a = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, self.t, 4)
        out = np.empty_like(a)
        self._check(self._lib.imt_poseidon_permute(self._h, _ptr(a), a.shape[0], _ptr(out)))
        return out

    def hash2_dev(self, d_in, n, d_out):
        self._check(self._lib.imt_poseidon_hash2_dev(self._h, _dev_ptr(d_in), n, _dev_ptr(d_out)))

    def hash3_dev(self, d_in, n, d_out):
        self._check(self._lib.imt_poseidon_hash3_dev(self._h, _dev_ptr(d_in), n, _dev_ptr(d_out)))

    def trace_hashes(self, inputs, arity, want_states=True):
        a = np.ascontiguousarray(inputs, dtype=np.uint64).reshape(-1, arity, 4)
        n = a.shape[0]
        states = np.empty((n, self.states_per_hash(arity), self.t, 4), np.uint64) if want_states else None
        digests = np.empty((n, 4), np.uint64)
        self._check(self._lib.imt_trace_hashes(self._h, _ptr(a), arity, n, _ptr(states), _ptr(digests)))
        return digests, states

    def sbox_cells_per_hash(self, arity):
        """S-boxes of one hash of `arity` inputs (each traced as 3 FE: x^2, x^4, x^5 + c) in the extended trace"""
        return (arity // self.rate + 1) * (self.r_f * self.t + self.r_p)

    def trace_hashes_ext(self, inputs, arity):
        """trace_hashes + the extended S-box trace: (digests, states, sbox[n, S-boxes per hash, 3, 4])"""
        a = np.ascontiguousarray(inputs, dtype=np.uint64).reshape(-1, arity, 4)
        n = a.shape[0]
        states = np.empty((n, self.states_per_hash(arity), self.t, 4), np.uint64)
        sbox = np.empty((n, self.sbox_cells_per_hash(arity), 3, 4), np.uint64)
        digests = np.empty((n, 4), np.uint64)
        self._check(self._lib.imt_poseidon_trace_ext(self._h, _ptr(a), arity, n, _ptr(states), _ptr(sbox), _ptr(digests)))
        return digests, states, sbox

    def trace_hashes_dev(self, d_in, arity, n, d_states, d_digests):
        self._check(self._lib.imt_trace_hashes_dev(self._h, _dev_ptr(d_in), arity, n,
                                                   _dev_ptr(d_states) if d_states is not None else None,
                                                   _dev_ptr(d_digests) if d_digests is not None else None))

    # ---- trees
    def _tree(self, fn, buf, n):
        h = ctypes.c_void_p()
        self._check(fn(self._h, buf, n, ctypes.byref(h)))
        return Tree(self, h)

    def build_from_hashes(self, leaf_hashes):
        a = _fe_array(leaf_hashes, ())
        return self._tree(self._lib.imt_tree_build_from_hashes, _ptr(a), a.shape[0])

    def build_from_leaves(self, preimages):
        a = _fe_array(preimages, (3,))
        return self._tree(self._lib.imt_tree_build_from_leaves, _ptr(a), a.shape[0])

    def build_from_leaves_ptr(self, host_ptr, n):
        """host pointer (e.g. a pinned torch tensor's data_ptr()) to n x 3 FE"""
        return self._tree(self._lib.imt_tree_build_from_leaves, ctypes.c_void_p(host_ptr), n)

    def build_from_leaves_dev(self, d_preimages, n):
        return self._tree(self._lib.imt_tree_build_from_leaves_dev, _dev_ptr(d_preimages), n)

    def build_from_hashes_dev(self, d_hashes, n):
        return self._tree(self._lib.imt_tree_build_from_hashes_dev, _dev_ptr(d_hashes), n)

    def load_tree(self, path):
        """imt_tree_load: rebuilds a tree from a checkpoint written by Tree.save() (flat header + n x 96 bytes of canonical
        little-endian `val, next_val, next_idx` + root): the leaves are re-hashed on the GPU and the file is refused
        (ImtError) when the rebuilt root differs from the stored one."""
        import os
        h = ctypes.c_void_p()
        self._check(self._lib.imt_tree_load(self._h, os.fsencode(path), ctypes.byref(h)))
        return Tree(self, h)

    @staticmethod
    def checkpoint_info(path):
        """header + root of a checkpoint file (no device needed): dict(num_leaves, depth, instance, root)"""
        import os
        info = _ffi.CheckpointInfo()
        st = _ffi.load().imt_checkpoint_read_info(os.fsencode(path), ctypes.byref(info))
        if st != _ffi.OK:
            raise ImtError(st, f"{path} is not an imt_b200 checkpoint")
        return dict(num_leaves=int(info.num_leaves), depth=int(info.depth), instance=(info.t, info.rate, info.r_f, info.r_p),
                    root=np.frombuffer(bytes(info.root), dtype=np.uint64).copy())

    # ---- path folding
    def verify_proofs(self, leaves, indices, roots, siblings):
        lv = _fe_array(leaves, ())
        q = lv.shape[0]
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(q)
        sib = _fe_array(siblings, ()).reshape(q, -1, 4) if q else np.zeros((0, 0, 4), np.uint64)
        depth = sib.shape[1]
        rt = _fe_array(roots, ())
        if rt.shape[0] == 1 and q != 1:
            rt = np.ascontiguousarray(np.broadcast_to(rt, (q, 4)))
        ok = np.zeros(q, np.uint8)
        self._check(self._lib.imt_verify_proofs(self._h, _ptr(lv), _ptr(idx), _ptr(rt), _ptr(sib), q, depth, _ptr(ok)))
        return ok.astype(bool)

    def trace_merkle_proofs(self, leaves, indices, siblings, want_states=True, out_states=None):
        """out_states: optional preallocated (q, depth, 132, 3, 4) uint64 array to receive the traces — e.g. a view of
        pinned host memory, which the device can fill at PCIe speed (the traces are 12 672 bytes per hash)."""
        lv = _fe_array(leaves, ())
        q = lv.shape[0]
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(q)
        sib = _fe_array(siblings, ()).reshape(q, -1, 4)
        depth = sib.shape[1]
        shape = (q, depth, self.states_per_hash(2), self.t, 4)
        if out_states is not None:
            if out_states.shape != shape or out_states.dtype != np.uint64 or not out_states.flags.c_contiguous:
                raise ValueError(f"out_states must be a C-contiguous uint64 array of shape {shape}")
            states = out_states
        else:
            states = np.empty(shape, np.uint64) if want_states else None
        roots = np.empty((q, 4), np.uint64)
        self._check(self._lib.imt_trace_merkle_proofs(self._h, _ptr(lv), _ptr(idx), _ptr(sib), q, depth, _ptr(states), _ptr(roots)))
        return roots, states

    def verify_proofs_dev(self, d_leaves, d_indices, d_roots, d_siblings, q, depth, d_ok):
        self._check(self._lib.imt_verify_proofs_dev(self._h, _dev_ptr(d_leaves), _dev_ptr(d_indices), _dev_ptr(d_roots), _dev_ptr(d_siblings),
                                                    q, depth, _dev_ptr(d_ok)))

    def trace_merkle_proofs_dev(self, d_leaves, d_indices, d_siblings, q, depth, d_states, d_roots):
        self._check(self._lib.imt_trace_merkle_proofs_dev(self._h, _dev_ptr(d_leaves), _dev_ptr(d_indices), _dev_ptr(d_siblings), q, depth,
                                                          _dev_ptr(d_states) if d_states is not None else None,
                                                          _dev_ptr(d_roots) if d_roots is not None else None))

    def trace_insert_witness(self, w, first_idx, out_states=None):
        """imt_insert_witness_trace — the Poseidon witness traces of everything the chip's insert_leaf hashes, per insert and in
        its call order (indexed_merkle_tree.rs:253-313): 3 leaf hashes + 4 x depth node hashes = (6 + 8 depth) permutations, ONE
        C call for the whole batch. `w` is the dict Tree.insert_batch returned, first_idx the slot of its first insert.
        Returns views into states[b][3 + 4 depth][132][3][4] under the names of the chip's steps, plus the roots each fold
        ends in, the rewired low leaf and the 128-bit limb witnesses."""
        b, d = w["low_idx"].shape[0], w["low_siblings"].shape[1]
        S = 3 + 4 * d
        shape = (b, S, self.states_per_hash(2), self.t, 4)
        states = out_states if out_states is not None else np.empty(shape, np.uint64)
        if states.shape != shape or states.dtype != np.uint64 or not states.flags.c_contiguous:
            raise ValueError(f"out_states must be a C-contiguous uint64 array of shape {shape}")
        roots, new_low = np.empty((b, 4, 4), np.uint64), np.empty((b, 3, 4), np.uint64)
        limbs, flags = np.empty((b, 6, 4), np.uint64), np.empty((b, 3), np.uint8)
        cw = _ffi.InsertWitness()
        keep = [np.ascontiguousarray(w[k]) for k in ("low_idx", "low_leaves", "low_siblings", "new_leaves", "new_siblings")]
        if w.get("fold_nodes") is not None:  # the one-launch form
            keep.append(np.ascontiguousarray(w["fold_nodes"]))
        for k, a in zip(("low_idx", "low_leaves", "low_siblings", "new_leaves", "new_siblings", "fold_nodes"), keep):
            setattr(cw, k, a.ctypes.data)
        self._check(self._lib.imt_insert_witness_trace(self._h, ctypes.byref(cw), b, d, int(first_idx), _ptr(states), _ptr(roots), _ptr(new_low),
                                                       _ptr(limbs), _ptr(flags)))
        return dict(states=states, limbs=limbs, limb_flags=flags.astype(bool), low_leaf=states[:, 0], low_path=states[:, 1:1 + d],
                    new_low_leaf=states[:, d + 1], interim_path=states[:, d + 2:2 * d + 2], zero_path=states[:, 2 * d + 2:3 * d + 2],
                    new_leaf=states[:, 3 * d + 2], new_path=states[:, 3 * d + 3:4 * d + 3], old_root=roots[:, 0], interim_root=roots[:, 1],
                    zero_leaf_root=roots[:, 2], new_root=roots[:, 3], new_low_leaf_preimage=new_low)

    def trace_insert_witness_dev(self, d_w, b, depth, first_idx, d_states=None, d_roots=None, d_new_low=None, d_limbs=None, d_flags=None):
        """imt_insert_witness_trace_dev: everything on the device. d_w: dict with CUDA tensors low_leaves, low_idx, low_siblings,
        new_leaves, new_siblings (the layouts Tree.insert_batch returns); outputs are CUDA tensors or None."""
        cw = _ffi.InsertWitness()
        for k in ("low_idx", "low_leaves", "low_siblings", "new_leaves", "new_siblings", "fold_nodes"):
            if d_w.get(k) is not None:
                setattr(cw, k, _dev_ptr(d_w[k]).value)
        opt = lambda t: _dev_ptr(t) if t is not None else None
        self._check(self._lib.imt_insert_witness_trace_dev(self._h, ctypes.byref(cw), b, depth, int(first_idx), opt(d_states), opt(d_roots),
                                                           opt(d_new_low), opt(d_limbs), opt(d_flags)))

    def non_inclusion_limbs(self, low_leaves, new_vals):
        """128-bit limb witnesses of verify_non_inclusion (indexed_merkle_tree.rs:143-172, 206-222): (b, 6, 4) FE in the
        chip's load order nl_q, nl_r, ll_q, ll_r, llv_q, llv_r and (b, 3) flags [nl < ll, llv < nl, passes the assertions]"""
        lv = _fe_array(low_leaves, (3,))
        nv = _fe_array(new_vals, ())
        b = nv.shape[0]
        limbs, flags = np.empty((b, 6, 4), np.uint64), np.empty((b, 3), np.uint8)
        self._check(self._lib.imt_non_inclusion_limbs(self._h, _ptr(lv), _ptr(nv), b, _ptr(limbs), _ptr(flags)))
        return limbs, flags.astype(bool)

    def shard_insert_plan(self, values, first_idx, pred_keys, pred_slots, succ_keys, succ_slots, flags):
        """replicated plan of a sharded insert chunk from the gathered [world][b] neighbour arrays"""
        v = _fe_array(values, ())
        b = v.shape[0]
        pk = np.ascontiguousarray(pred_keys, dtype=np.uint64).reshape(-1, b, 4)
        world = pk.shape[0]
        sk = np.ascontiguousarray(succ_keys, dtype=np.uint64).reshape(world, b, 4)
        ps = np.ascontiguousarray(pred_slots, dtype=np.uint64).reshape(world, b)
        ss = np.ascontiguousarray(succ_slots, dtype=np.uint64).reshape(world, b)
        fl = np.ascontiguousarray(flags, dtype=np.uint8).reshape(world, b)
        x, upd = np.empty(2 * b, np.uint64), np.empty((2 * b, 3, 4), np.uint64)
        low_old, largest = np.empty((b, 3, 4), np.uint64), np.empty(b, np.uint8)
        self._check(self._lib.imt_shard_insert_plan(self._h, _ptr(v), b, int(first_idx), world, _ptr(pk), _ptr(ps), _ptr(sk), _ptr(ss), _ptr(fl),
                                                    _ptr(x), _ptr(upd), _ptr(low_old), _ptr(largest)))
        return x, upd, low_old, largest

    def low_leaf_merge(self, values, cand_keys, cand_slots, flags, occupied_total, n_total, head_next_zero):
        """replicated half of a sharded lookup: [world][q] gathered candidates -> (low_idx, matched)"""
        v = _fe_array(values, ())
        q = v.shape[0]
        ck = np.ascontiguousarray(cand_keys, dtype=np.uint64).reshape(-1, q, 4)
        world = ck.shape[0]
        cs = np.ascontiguousarray(cand_slots, dtype=np.uint64).reshape(world, q)
        fl = np.ascontiguousarray(flags, dtype=np.uint8).reshape(world, q)
        low, matched = np.empty(q, np.uint64), np.empty(q, np.uint8)
        self._check(self._lib.imt_low_leaf_merge(self._h, _ptr(v), _ptr(ck), _ptr(cs), _ptr(fl), world, q, int(occupied_total),
                                                 int(n_total), 1 if head_next_zero else 0, _ptr(low), _ptr(matched)))
        return low, matched.astype(bool)

    # ---- multi-GPU inside the library (one process per GPU): NCCL communicator attached to this context
    @staticmethod
    def comm_unique_id():
        """128 bytes from ncclGetUniqueId: rank 0 makes it and hands it to every rank (any transport)"""
        lib = _ffi.load()
        buf = (ctypes.c_uint8 * 128)()
        st = lib.imt_comm_unique_id(buf)
        if st != _ffi.OK:
            raise ImtError(st, "imt_comm_unique_id failed: libnccl.so.2 could not be loaded")
        return bytes(buf)

    def comm_create(self, rank, world, unique_id):
        uid = (ctypes.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._check(self._lib.imt_comm_create(self._h, int(rank), int(world), uid))

    def comm_destroy(self):
        self._check(self._lib.imt_comm_destroy(self._h))

    def comm_info(self):
        """(rank, world, NCCL version code) of this context's group"""
        r, w, v = ctypes.c_uint(), ctypes.c_uint(), ctypes.c_int()
        self._check(self._lib.imt_comm_info(self._h, ctypes.byref(r), ctypes.byref(w), ctypes.byref(v)))
        return r.value, w.value, v.value

    def sharded_build_from_leaves(self, local_preimages):
        """IndexedMerkleTree::new for this rank's slice of the leaves, root exchange (NCCL, inside the library) included"""
        a = _fe_array(local_preimages, (3,))
        return self._tree(self._lib.imt_sharded_build_from_leaves, _ptr(a), a.shape[0])

    def sharded_build_from_leaves_dev(self, d_local_preimages, n_local):
        return self._tree(self._lib.imt_sharded_build_from_leaves_dev, _dev_ptr(d_local_preimages), n_local)

    def calibrate_imad(self, ms=200.0):
        rate, mhz = ctypes.c_double(), ctypes.c_double()
        self._check(self._lib.imt_calibrate_imad(self._h, float(ms), ctypes.byref(rate), ctypes.byref(mhz)))
        return rate.value, mhz.value


class Tree:
    """imt_tree: every level of the tree, resident on the GPU (utils.rs:5-10's `tree: Vec<Vec<F>>`)."""

    def __init__(self, engine, handle):
        self.engine = engine
        self._lib = engine._lib
        self._h = handle
        engine._trees.add(self)

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self.engine, "_h", None):  # the context is gone => Engine.close() already destroyed this tree
                self._lib.imt_tree_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_leaves(self):
        return int(self._lib.imt_tree_num_leaves(self._h))

    @property
    def depth(self):
        return int(self._lib.imt_tree_depth(self._h))

    def root(self):
        out = np.empty(4, np.uint64)
        self.engine._check(self._lib.imt_tree_root(self._h, _ptr(out)))
        return out

    def subtree_root(self):
        """this rank's OWN root (level `local depth` of its subtree) — root() reports the global root once a cap is attached"""
        n_local = self.shard_info()[2]
        return self.level(n_local.bit_length() - 1, 1)[0]

    def shard_info(self):
        """(rank, world, leaves held by this rank)"""
        r, w, n = ctypes.c_uint(), ctypes.c_uint(), ctypes.c_size_t()
        self.engine._check(self._lib.imt_tree_shard_info(self._h, ctypes.byref(r), ctypes.byref(w), ctypes.byref(n)))
        return int(r.value), int(w.value), int(n.value)

    def root_dev(self, d_out):
        self.engine._check(self._lib.imt_tree_root_dev(self._h, _dev_ptr(d_out)))

    def attach_cap_dev(self, rank, world, d_roots):
        self.engine._check(self._lib.imt_tree_attach_cap_dev(self._h, rank, world, _dev_ptr(d_roots)))

    def level(self, lvl, count):
        out = np.empty((count, 4), np.uint64)
        self.engine._check(self._lib.imt_tree_level(self._h, lvl, _ptr(out)))
        return out

    def preimages(self, n_local):
        out = np.empty((n_local, 3, 4), np.uint64)
        self.engine._check(self._lib.imt_tree_preimages(self._h, _ptr(out)))
        return out

    def save(self, path):
        """imt_tree_save — checkpoint: a 64-byte header, the leaves as the reference's serde derive would lay
        them out (utils.rs:12-17: val, next_val, next_idx, each the canonical 32-byte little-endian `to_repr` bytes whatever
        the engine's format) and the canonical root. The levels are NOT stored: load_tree() re-hashes (0.55 s at depth 24)
        and verifies the root."""
        import os
        self.engine._check(self._lib.imt_tree_save(self._h, os.fsencode(path)))

    def rebuild_from_leaves(self, preimages):
        a = _fe_array(preimages, (3,))
        self.engine._check(self._lib.imt_tree_rebuild_from_leaves(self._h, _ptr(a)))

    def rebuild_from_leaves_ptr(self, host_ptr):
        """host pointer (e.g. a pinned torch tensor's data_ptr()) to n x 3 FE"""
        self.engine._check(self._lib.imt_tree_rebuild_from_leaves(self._h, ctypes.c_void_p(host_ptr)))

    def rebuild_from_leaves_dev(self, d_preimages):
        self.engine._check(self._lib.imt_tree_rebuild_from_leaves_dev(self._h, _dev_ptr(d_preimages)))

    def get_proofs(self, indices, helpers_as_fe=False):
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
        q, d = idx.shape[0], self.depth
        sib = np.empty((q, d, 4), np.uint64)
        if helpers_as_fe:
            hel = np.empty((q, d, 4), np.uint64)
            self.engine._check(self._lib.imt_tree_get_proofs_fe(self._h, _ptr(idx), q, _ptr(sib), _ptr(hel)))
        else:
            hel = np.empty((q, d), np.uint8)
            self.engine._check(self._lib.imt_tree_get_proofs(self._h, _ptr(idx), q, _ptr(sib), _ptr(hel)))
        return sib, hel

    def trace_proofs(self, indices, out_states=None):
        """witness traces of verify_merkle_proof for leaves of this tree, by index: (q, depth, states per hash, t, 4).
        Every traced hash reads stored nodes, so they all run in parallel (imt_tree_trace_proofs)."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
        q = idx.shape[0]
        shape = (q, self.depth, self.engine.states_per_hash(2), self.engine.t, 4)
        if out_states is not None:
            if out_states.shape != shape or out_states.dtype != np.uint64 or not out_states.flags.c_contiguous:
                raise ValueError(f"out_states must be a C-contiguous uint64 array of shape {shape}")
        states = out_states if out_states is not None else np.empty(shape, np.uint64)
        self.engine._check(self._lib.imt_tree_trace_proofs(self._h, _ptr(idx), q, _ptr(states)))
        return states

    def trace_proofs_ext(self, indices):
        """trace_proofs + the extended S-box trace: (states, sbox[q, depth, S-boxes per hash, 3, 4])"""
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
        q, e = idx.shape[0], self.engine
        states = np.empty((q, self.depth, e.states_per_hash(2), e.t, 4), np.uint64)
        sbox = np.empty((q, self.depth, e.sbox_cells_per_hash(2), 3, 4), np.uint64)
        e._check(self._lib.imt_tree_trace_proofs_ext(self._h, _ptr(idx), q, _ptr(states), _ptr(sbox)))
        return states, sbox

    def trace_proofs_dev(self, d_indices, q, d_states):
        self.engine._check(self._lib.imt_tree_trace_proofs_dev(self._h, _dev_ptr(d_indices), q, _dev_ptr(d_states)))

    def get_proofs_dev(self, d_indices, q, d_siblings, d_helpers=None):
        self.engine._check(self._lib.imt_tree_get_proofs_dev(self._h, _dev_ptr(d_indices), q, _dev_ptr(d_siblings),
                                                             _dev_ptr(d_helpers) if d_helpers is not None else None))

    def low_leaf_lookup_dev(self, d_values, q, d_low_idx, d_matched=None):
        self.engine._check(self._lib.imt_low_leaf_lookup_dev(self._h, _dev_ptr(d_values), q, _dev_ptr(d_low_idx),
                                                             _dev_ptr(d_matched) if d_matched is not None else None))

    @property
    def occupied(self):
        """number of occupied slots = the slot the next insert goes to"""
        m = ctypes.c_size_t()
        self.engine._check(self._lib.imt_tree_occupied(self._h, ctypes.byref(m)))
        return int(m.value)

    def low_leaf_lookup(self, values):
        v = _fe_array(values, ())
        q = v.shape[0]
        low = np.empty(q, np.uint64)
        matched = np.empty(q, np.uint8)
        self.engine._check(self._lib.imt_low_leaf_lookup(self._h, _ptr(v), q, _ptr(low), _ptr(matched)))
        return low, matched.astype(bool)

    @staticmethod
    def non_inclusion_buffers(q, depth, pinned=False):
        """output buffers of non_inclusion_paths for q queries; pinned=True allocates page-locked host memory (torch), which
        the device fills at PCIe speed and which can be reused across calls (fresh pageable arrays cost a page fault per 4 KB)"""
        shapes = dict(low_idx=((q,), np.uint64), matched=((q,), np.uint8), low_leaves=((q, 3, 4), np.uint64),
                      siblings=((q, depth, 4), np.uint64), helpers=((q, depth), np.uint8), is_largest=((q,), np.uint8))
        if not pinned:
            return {k: np.empty(sh, dt) for k, (sh, dt) in shapes.items()}
        import torch
        keep = {k: torch.empty(sh, dtype=torch.int64 if dt == np.uint64 else torch.uint8, pin_memory=True) for k, (sh, dt) in shapes.items()}
        out = {k: (t.numpy().view(np.uint64) if shapes[k][1] == np.uint64 else t.numpy()) for k, t in keep.items()}
        out["_pinned"] = keep  # keeps the page-locked allocations alive
        return out

    def non_inclusion_paths(self, values, out=None):
        """low leaf + its Merkle path + flags for every value (verify_non_inclusion's witnesses, IMT:127-137).
        out: buffers from non_inclusion_buffers() to fill instead of allocating new arrays."""
        v = _fe_array(values, ())
        q, d = v.shape[0], self.depth
        o = out if out is not None else self.non_inclusion_buffers(q, d)
        if o["siblings"].shape != (q, d, 4) or o["low_idx"].shape != (q,):
            raise ValueError("out buffers were made for another batch size / depth")
        self.engine._check(self._lib.imt_non_inclusion_paths(self._h, _ptr(v), q, _ptr(o["low_idx"]), _ptr(o["matched"]),
                                                             _ptr(o["low_leaves"]), _ptr(o["siblings"]), _ptr(o["helpers"]),
                                                             _ptr(o["is_largest"])))
        return o

    def trace_non_inclusion(self, values, out=None, out_states=None, states=True):
        """imt_non_inclusion_witness_trace — the whole witness of the chip's verify_non_inclusion (indexed_merkle_tree.rs:127-229) for
        every value in ONE call: the lookup, the low leaf + path + flags (as non_inclusion_paths), the 128-bit limb witnesses and the
        Poseidon states of the 1 + depth hashes it constrains ([0] H3(low leaf), [1..depth] its fold up the path), shape
        (q, 1 + depth, 132, 3, 4). Returns the non_inclusion_paths dict + limbs, limb_flags, states, leaf_hash (view), path (view)."""
        v = _fe_array(values, ())
        q, d = v.shape[0], self.depth
        o = out if out is not None else self.non_inclusion_buffers(q, d)
        if o["siblings"].shape != (q, d, 4) or o["low_idx"].shape != (q,):
            raise ValueError("out buffers were made for another batch size / depth")
        o = dict(o)
        o["limbs"], flags = np.empty((q, 6, 4), np.uint64), np.empty((q, 3), np.uint8)
        st = None
        if states:
            shape = (q, 1 + d, self.engine.states_per_hash(2), self.engine.t, 4)
            st = out_states if out_states is not None else np.empty(shape, np.uint64)
            if st.shape != shape or st.dtype != np.uint64 or not st.flags.c_contiguous:
                raise ValueError(f"out_states must be a C-contiguous uint64 array of shape {shape}")
        self.engine._check(self._lib.imt_non_inclusion_witness_trace(self._h, _ptr(v), q, _ptr(o["low_idx"]), _ptr(o["matched"]), _ptr(o["low_leaves"]),
                                                                     _ptr(o["siblings"]), _ptr(o["helpers"]), _ptr(o["is_largest"]), _ptr(o["limbs"]),
                                                                     _ptr(flags), _ptr(st)))
        o["limb_flags"] = flags.astype(bool)
        if st is not None:
            o.update(states=st, leaf_hash=st[:, 0], path=st[:, 1:])
        return o

    def trace_non_inclusion_dev(self, d_values, q, d_low_idx, d_low_leaves, d_states, d_matched=None, d_siblings=None, d_helpers=None,
                                d_is_largest=None, d_limbs=None, d_limb_flags=None):
        """imt_non_inclusion_witness_trace_dev: CUDA tensors in and out (d_low_idx and d_low_leaves are required, the rest may be None)"""
        opt = lambda t: _dev_ptr(t) if t is not None else None
        self.engine._check(self._lib.imt_non_inclusion_witness_trace_dev(self._h, _dev_ptr(d_values), q, _dev_ptr(d_low_idx), opt(d_matched),
                                                                         _dev_ptr(d_low_leaves), opt(d_siblings), opt(d_helpers), opt(d_is_largest),
                                                                         opt(d_limbs), opt(d_limb_flags), opt(d_states)))

    @staticmethod
    def insert_buffers(b, depth, pinned=False, fold_nodes=True):
        """output buffers of insert_batch for b inserts (see non_inclusion_buffers for `pinned`). fold_nodes: also the chain values of
        the four folds insert_leaf constrains (b, 4, depth) FE, which make trace_insert_witness a single launch (single-GPU trees)."""
        shapes = dict(old_roots=((b, 4), np.uint64), low_idx=((b,), np.uint64), low_leaves=((b, 3, 4), np.uint64),
                      low_siblings=((b, depth, 4), np.uint64), low_helpers=((b, depth), np.uint8), new_roots=((b, 4), np.uint64),
                      new_leaves=((b, 3, 4), np.uint64), new_siblings=((b, depth, 4), np.uint64), new_helpers=((b, depth), np.uint8),
                      is_largest=((b,), np.uint8))
        if fold_nodes:
            shapes["fold_nodes"] = ((b, 4, depth, 4), np.uint64)
        if not pinned:
            return {k: np.empty(sh, dt) for k, (sh, dt) in shapes.items()}
        import torch
        keep = {k: torch.empty(sh, dtype=torch.int64 if dt == np.uint64 else torch.uint8, pin_memory=True) for k, (sh, dt) in shapes.items()}
        out = {k: (t.numpy().view(np.uint64) if shapes[k][1] == np.uint64 else t.numpy()) for k, t in keep.items()}
        out["_pinned"] = keep
        return out

    def insert_batch(self, new_vals, first_idx=None, out=None):
        """imt_insert_batch: the tree advances by the whole batch; returns every per-insert witness insert_leaf loads
        (IMT:444-489). out: buffers from insert_buffers() to fill instead of allocating new arrays."""
        v = _fe_array(new_vals, ())
        if first_idx is None:
            first_idx = self.occupied
        b, d = v.shape[0], self.depth
        o = out if out is not None else self.insert_buffers(b, d)
        if o["low_siblings"].shape != (b, d, 4):
            raise ValueError("out buffers were made for another batch size / depth")
        w = _ffi.InsertWitness(*[o[k].ctypes.data if k in o else None for k, _ in _ffi.InsertWitness._fields_])
        self.engine._check(self._lib.imt_insert_batch(self._h, _ptr(v), b, int(first_idx), ctypes.byref(w)))
        return o

    # ---- sharded tree, exchanges inside the library (the engine has a communicator: Engine.comm_create)
    def exchange_roots(self):
        """NCCL all-gather of the subtree roots out of / into the trees' device buffers + the cap levels"""
        self.engine._check(self._lib.imt_tree_exchange_roots(self._h))

    def sharded_rebuild_from_leaves(self, local_preimages):
        a = _fe_array(local_preimages, (3,))
        self.engine._check(self._lib.imt_sharded_rebuild_from_leaves(self._h, _ptr(a)))

    def sharded_rebuild_from_leaves_ptr(self, host_ptr):
        self.engine._check(self._lib.imt_sharded_rebuild_from_leaves(self._h, ctypes.c_void_p(host_ptr)))

    def sharded_rebuild_from_leaves_dev(self, d_local_preimages):
        self.engine._check(self._lib.imt_sharded_rebuild_from_leaves_dev(self._h, _dev_ptr(d_local_preimages)))

    def sharded_get_proofs(self, indices):
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
        q, d = idx.shape[0], self.depth
        sib, hel = np.empty((q, d, 4), np.uint64), np.empty((q, d), np.uint8)
        self.engine._check(self._lib.imt_sharded_get_proofs(self._h, _ptr(idx), q, _ptr(sib), _ptr(hel)))
        return sib, hel

    def sharded_leaves(self, indices):
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
        q = idx.shape[0]
        out, lg = np.empty((q, 3, 4), np.uint64), np.empty(q, np.uint8)
        self.engine._check(self._lib.imt_sharded_leaves(self._h, _ptr(idx), q, _ptr(out), _ptr(lg)))
        return out, lg

    def sharded_low_leaf_lookup(self, values):
        v = _fe_array(values, ())
        q = v.shape[0]
        low, matched = np.empty(q, np.uint64), np.empty(q, np.uint8)
        self.engine._check(self._lib.imt_sharded_low_leaf_lookup(self._h, _ptr(v), q, _ptr(low), _ptr(matched)))
        return low, matched.astype(bool)

    @property
    def sharded_occupied(self):
        m = ctypes.c_uint64()
        self.engine._check(self._lib.imt_sharded_occupied(self._h, ctypes.byref(m)))
        return int(m.value)

    def sharded_non_inclusion_paths(self, values, out=None):
        v = _fe_array(values, ())
        q, d = v.shape[0], self.depth
        o = out if out is not None else self.non_inclusion_buffers(q, d)
        self.engine._check(self._lib.imt_sharded_non_inclusion_paths(self._h, _ptr(v), q, _ptr(o["low_idx"]), _ptr(o["matched"]),
                                                                     _ptr(o["low_leaves"]), _ptr(o["siblings"]), _ptr(o["helpers"]),
                                                                     _ptr(o["is_largest"])))
        return o

    def sharded_insert_batch(self, new_vals, first_idx=None, out=None):
        v = _fe_array(new_vals, ())
        if first_idx is None:
            first_idx = self.sharded_occupied
        b, d = v.shape[0], self.depth
        o = out if out is not None else self.insert_buffers(b, d, fold_nodes=False)
        w = _ffi.InsertWitness(*[o[k].ctypes.data if k in o and k != "fold_nodes" else None for k, _ in _ffi.InsertWitness._fields_])
        self.engine._check(self._lib.imt_sharded_insert_batch(self._h, _ptr(v), b, int(first_idx), ctypes.byref(w)))
        return o

    def sharded_trace_proofs(self, indices, out_states=None):
        """owner-sharded witness traces: (positions into `indices` of the queries this rank owns, their traces)"""
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
        q = idx.shape[0]
        pos, n_mine = np.empty(q, np.uint64), ctypes.c_size_t()
        if out_states is None:  # sized for the queries this rank owns
            lo = np.uint64(self.shard_rank * self.shard_leaves)
            owned = int(np.count_nonzero((idx >= lo) & (idx < lo + np.uint64(self.shard_leaves))))
            out_states = np.empty((owned, self.depth, self.engine.states_per_hash(2), self.engine.t, 4), np.uint64)
        self.engine._check(self._lib.imt_sharded_trace_proofs(self._h, _ptr(idx), q, _ptr(pos), ctypes.byref(n_mine), _ptr(out_states)))
        return pos[: n_mine.value].astype(np.int64), out_states[: n_mine.value]

    @property
    def shard_rank(self):
        return self.shard_info()[0]

    @property
    def shard_leaves(self):
        """leaves held by this rank"""
        return self.shard_info()[2]

    # ---- subtree sharding
    def set_shard(self, rank, world):
        self.engine._check(self._lib.imt_tree_set_shard(self._h, rank, world))

    @property
    def head_next_zero(self):
        f = ctypes.c_int()
        self.engine._check(self._lib.imt_tree_head_next_zero(self._h, ctypes.byref(f)))
        return bool(f.value)

    def low_leaf_candidates(self, values):
        v = _fe_array(values, ())
        q = v.shape[0]
        keys, slots, flags = np.empty((q, 4), np.uint64), np.empty(q, np.uint64), np.empty(q, np.uint8)
        self.engine._check(self._lib.imt_low_leaf_candidates(self._h, _ptr(v), q, _ptr(keys), _ptr(slots), _ptr(flags)))
        return keys, slots, flags

    def shard_insert_neighbors(self, values):
        v = _fe_array(values, ())
        b = v.shape[0]
        pk, sk = np.empty((b, 4), np.uint64), np.empty((b, 4), np.uint64)
        ps, ss, fl = np.empty(b, np.uint64), np.empty(b, np.uint64), np.empty(b, np.uint8)
        self.engine._check(self._lib.imt_shard_insert_neighbors(self._h, _ptr(v), b, _ptr(pk), _ptr(ps), _ptr(sk), _ptr(ss), _ptr(fl)))
        return pk, ps, sk, ss, fl

    def shard_insert_apply(self, x, upd, local_depth):
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1)
        u = _fe_array(upd, (3,))
        b = x.shape[0] // 2
        sub_roots, sib = np.empty((2 * b, 4), np.uint64), np.empty((2 * b, local_depth, 4), np.uint64)
        self.engine._check(self._lib.imt_shard_insert_apply(self._h, _ptr(x), _ptr(u), b, _ptr(sub_roots), _ptr(sib)))
        return sub_roots, sib

    def shard_insert_cap(self, x, sub_roots, cap_depth):
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1)
        r = _fe_array(sub_roots, ())
        b = x.shape[0] // 2
        roots, sib = np.empty((2 * b, 4), np.uint64), np.empty((2 * b, cap_depth, 4), np.uint64)
        self.engine._check(self._lib.imt_shard_insert_cap(self._h, _ptr(x), _ptr(r), b, _ptr(roots), _ptr(sib)))
        return roots, sib

    def leaves(self, indices):
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
        q = idx.shape[0]
        out, lg = np.empty((q, 3, 4), np.uint64), np.empty(q, np.uint8)
        self.engine._check(self._lib.imt_tree_leaves(self._h, _ptr(idx), q, _ptr(out), _ptr(lg)))
        return out, lg

    def attach_cap(self, rank, world, subtree_roots):
        """after (re)building: exchange the N subtree roots (root() of each rank) and attach them here"""
        r = _fe_array(subtree_roots, ())
        if r.shape[0] != world:
            raise ValueError("need one subtree root per rank")
        self.engine._check(self._lib.imt_tree_attach_cap(self._h, rank, world, _ptr(r)))


class Multi:
    """imt_multi: ONE process drives N GPUs (a power of two). One context per device + their NCCL communicators
    (ncclCommInitAll) inside the library; build_from_leaves() is `IndexedMerkleTree::new` over all of them in one call."""

    def __init__(self, devices, fmt="canonical"):
        self._lib = _ffi.load()
        self.fmt = {"canonical": _ffi.FE_CANONICAL, "montgomery": _ffi.FE_MONTGOMERY}[fmt]
        devs = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        h = ctypes.c_void_p()
        st = self._lib.imt_multi_create(devs, len(devices), self.fmt, ctypes.byref(h))
        if st != _ffi.OK:
            raise ImtError(st, f"imt_multi_create({list(devices)}) failed with status {st}")
        self._h = h
        self.devices = [int(d) for d in devices]
        self._trees = weakref.WeakSet()

    @property
    def size(self):
        return int(self._lib.imt_multi_size(self._h))

    @property
    def nccl_version(self):
        """NCCL version code in use; 0 = the copy transport (a device listed twice)"""
        return int(self._lib.imt_multi_uses_nccl(self._h))

    def engine(self, i):
        """Engine view of device i's context (owned by this Multi)"""
        e = Engine.__new__(Engine)
        e._lib, e.fmt, e._h = self._lib, self.fmt, ctypes.c_void_p(self._lib.imt_multi_ctx(self._h, i))
        e.t, e.rate, e.r_f, e.r_p, e.generic, e.device = 3, 2, 8, 57, False, self.devices[i]
        e.states_per_perm = 66
        e._trees = weakref.WeakSet()
        e._borrowed = True
        return e

    def _check(self, st):
        if st != _ffi.OK:
            msg = self._lib.imt_multi_last_error(self._h).decode() or self._lib.imt_status_string(st).decode()
            raise ImtError(st, msg)

    def build_from_leaves(self, preimages):
        a = _fe_array(preimages, (3,))
        h = ctypes.c_void_p()
        self._check(self._lib.imt_multi_build_from_leaves(self._h, _ptr(a), a.shape[0], ctypes.byref(h)))
        return MTree(self, h)

    def build_from_leaves_ptr(self, host_ptr, n):
        h = ctypes.c_void_p()
        self._check(self._lib.imt_multi_build_from_leaves(self._h, ctypes.c_void_p(host_ptr), n, ctypes.byref(h)))
        return MTree(self, h)

    def load_tree(self, path):
        """imt_multi_load: a checkpoint (of a single-GPU or a sharded tree: the file is the same) into a tree sharded over this group"""
        import os
        h = ctypes.c_void_p()
        self._check(self._lib.imt_multi_load(self._h, os.fsencode(path), ctypes.byref(h)))
        return MTree(self, h)

    def close(self):
        if getattr(self, "_h", None):
            for t in list(self._trees):
                t.close()
            self._lib.imt_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MTree:
    """imt_mtree: a tree sharded by subtree over the devices of a Multi; host arrays are whole-tree arrays"""

    def __init__(self, multi, handle):
        self.multi, self._lib, self._h = multi, multi._lib, handle
        multi._trees.add(self)

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self.multi, "_h", None):
                self._lib.imt_mtree_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_leaves(self):
        return int(self._lib.imt_mtree_num_leaves(self._h))

    @property
    def depth(self):
        return int(self._lib.imt_mtree_depth(self._h))

    def root(self):
        out = np.empty(4, np.uint64)
        self.multi._check(self._lib.imt_mtree_root(self._h, _ptr(out)))
        return out

    def rebuild_from_leaves(self, preimages):
        a = _fe_array(preimages, (3,))
        self.multi._check(self._lib.imt_mtree_rebuild_from_leaves(self._h, _ptr(a)))

    def rebuild_from_leaves_ptr(self, host_ptr):
        self.multi._check(self._lib.imt_mtree_rebuild_from_leaves(self._h, ctypes.c_void_p(host_ptr)))

    def save(self, path):
        import os
        self.multi._check(self._lib.imt_mtree_save(self._h, os.fsencode(path)))

    def get_proofs(self, indices):
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
        q, d = idx.shape[0], self.depth
        sib, hel = np.empty((q, d, 4), np.uint64), np.empty((q, d), np.uint8)
        self.multi._check(self._lib.imt_mtree_get_proofs(self._h, _ptr(idx), q, _ptr(sib), _ptr(hel)))
        return sib, hel

    def leaves(self, indices):
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
        q = idx.shape[0]
        out, lg = np.empty((q, 3, 4), np.uint64), np.empty(q, np.uint8)
        self.multi._check(self._lib.imt_mtree_leaves(self._h, _ptr(idx), q, _ptr(out), _ptr(lg)))
        return out, lg

    def low_leaf_lookup(self, values):
        v = _fe_array(values, ())
        q = v.shape[0]
        low, matched = np.empty(q, np.uint64), np.empty(q, np.uint8)
        self.multi._check(self._lib.imt_mtree_low_leaf_lookup(self._h, _ptr(v), q, _ptr(low), _ptr(matched)))
        return low, matched.astype(bool)

    @property
    def occupied(self):
        m = ctypes.c_uint64()
        self.multi._check(self._lib.imt_mtree_occupied(self._h, ctypes.byref(m)))
        return int(m.value)

    def non_inclusion_paths(self, values, out=None):
        v = _fe_array(values, ())
        q, d = v.shape[0], self.depth
        o = out if out is not None else Tree.non_inclusion_buffers(q, d)
        self.multi._check(self._lib.imt_mtree_non_inclusion_paths(self._h, _ptr(v), q, _ptr(o["low_idx"]), _ptr(o["matched"]),
                                                                  _ptr(o["low_leaves"]), _ptr(o["siblings"]), _ptr(o["helpers"]),
                                                                  _ptr(o["is_largest"])))
        return o

    def insert_batch(self, new_vals, first_idx=None, out=None):
        v = _fe_array(new_vals, ())
        if first_idx is None:
            first_idx = self.occupied
        b, d = v.shape[0], self.depth
        o = out if out is not None else Tree.insert_buffers(b, d, fold_nodes=False)
        w = _ffi.InsertWitness(*[o[k].ctypes.data if k in o and k != "fold_nodes" else None for k, _ in _ffi.InsertWitness._fields_])
        self.multi._check(self._lib.imt_mtree_insert_batch(self._h, _ptr(v), b, int(first_idx), ctypes.byref(w)))
        return o

    def trace_proofs(self, indices, out_states=None):
        idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
        q = idx.shape[0]
        shape = (q, self.depth, STATES_PER_HASH, 3, 4)
        states = out_states if out_states is not None else np.empty(shape, np.uint64)
        if states.shape != shape or states.dtype != np.uint64 or not states.flags.c_contiguous:
            raise ValueError(f"out_states must be a C-contiguous uint64 array of shape {shape}")
        self.multi._check(self._lib.imt_mtree_trace_proofs(self._h, _ptr(idx), q, _ptr(states)))
        return states
