"""TEST DOUBLE (test infrastructure, never shipped): the subset of Engine / Tree that sharding.ShardedTree drives,
computed by the CPU oracle. It lets the world_size-2 gloo tests exercise the host-side sharding logic — root exchange,
owner routing, candidate merge — in a container without a GPU. The product path never imports this file."""
import bisect

import numpy as np

import oracle as O


class OracleTree:
    def __init__(self, pre):
        self.pre = np.array(pre, dtype=np.uint64).reshape(-1, 3, 4)  # a private copy: the insert steps update it in place
        self.n = self.pre.shape[0]
        self.local_depth = self.n.bit_length() - 1
        self.rank, self.world, self.cap = 0, 1, None
        self._build()

    def _build(self):
        self.tree = O.tree_build(O.hash3(self.pre, 4), 4)

    def rebuild_from_leaves(self, pre):
        self.pre = np.array(pre, dtype=np.uint64).reshape(-1, 3, 4)
        self.cap = None
        self._build()

    def set_shard(self, rank, world):
        self.rank, self.world = rank, world

    def root(self):
        return (self.cap if self.cap is not None else self.tree)[-1].copy()

    def subtree_root(self):
        return self.tree[-1].copy()

    def attach_cap(self, rank, world, roots):
        self.rank, self.world = rank, world
        self.cap = O.tree_build(np.ascontiguousarray(roots, dtype=np.uint64).reshape(world, 4), 1) if world > 1 else None

    @property
    def depth(self):
        return self.local_depth + (self.world.bit_length() - 1 if self.cap is not None else 0)

    def get_proofs(self, indices):
        idx = np.asarray(indices, dtype=np.uint64).reshape(-1)
        sib, hel = np.zeros((idx.size, self.depth, 4), np.uint64), np.zeros((idx.size, self.depth), np.uint8)
        base = self.rank * self.n if self.cap is not None else 0
        for k, g in enumerate(idx):
            local = int(g) - base
            if not 0 <= local < self.n:
                raise IndexError("index out of bounds")
            s, h = O.get_proof(self.tree, self.n, local)
            sib[k, : self.local_depth], hel[k, : self.local_depth] = s, h
            if self.cap is not None:
                s, h = O.get_proof(self.cap, self.world, self.rank)
                sib[k, self.local_depth:], hel[k, self.local_depth:] = s, h
        return sib, hel

    def trace_proofs(self, indices):
        """traces of the paths of this (sub)tree: every hash of the fold, operands read from the stored levels"""
        idx = np.asarray(indices, dtype=np.uint64).reshape(-1)
        out = np.zeros((idx.size, self.depth, 132, 3, 4), np.uint64)
        base = self.rank * self.n if self.cap is not None else 0
        lv_off = lambda n, l: 2 * n - 2 * (n >> l)  # noqa: E731  FE offset of level l in the concatenated tree
        for k, g in enumerate(idx):
            local = int(g) - base
            for l in range(self.depth):
                if l < self.local_depth:
                    tree, n, node = self.tree, self.n, (local >> l) & ~1
                    off = lv_off(n, l) + node
                else:
                    cl = l - self.local_depth
                    tree, n, node = self.cap, self.world, (self.rank >> cl) & ~1
                    off = lv_off(n, cl) + node
                flat = np.asarray(tree).reshape(-1, 4)
                out[k, l] = O.hash_trace(flat[off:off + 2])[1]
        return out

    def leaves(self, indices):
        idx = np.asarray(indices, dtype=np.int64).reshape(-1) - (self.rank * self.n if self.world > 1 else 0)
        out = self.pre[idx]
        return out, (~out[:, 1].any(axis=1)).astype(np.uint8)

    def _sorted(self):
        base = self.rank * self.n
        occ = [i for i in range(self.n) if base + i == 0 or self.pre[i, 0].any()]
        keys = sorted((O.to_int(self.pre[i, 0]), base + i) for i in occ)
        return keys

    @property
    def occupied(self):
        return len(self._sorted())

    @property
    def head_next_zero(self):
        return self.rank == 0 and not self.pre[0, 1].any()

    def low_leaf_candidates(self, values):
        v = O.to_ints(values)
        ks = self._sorted()
        only = [k for k, _ in ks]
        keys, slots, flags = np.zeros((len(v), 4), np.uint64), np.zeros(len(v), np.uint64), np.zeros(len(v), np.uint8)
        for i, x in enumerate(v):
            j = bisect.bisect_left(only, x)
            if j < len(only) and only[j] == x:
                flags[i] |= 2
            if j >= 1:
                keys[i], slots[i] = O.fe(ks[j - 1][0]), ks[j - 1][1]
                flags[i] |= 1
        return keys, slots, flags


    # ---- sharded inserts (the four per-chunk steps of include/imt_b200.h, restated with the oracle)
    def shard_insert_neighbors(self, values):
        v = O.to_ints(values)
        ks = self._sorted()
        only = [k for k, _ in ks]
        b = len(v)
        pk, sk = np.zeros((b, 4), np.uint64), np.zeros((b, 4), np.uint64)
        ps, ss, fl = np.zeros(b, np.uint64), np.zeros(b, np.uint64), np.zeros(b, np.uint8)
        for i, x in enumerate(v):
            j = bisect.bisect_left(only, x)
            js = j
            if j < len(only) and only[j] == x:
                fl[i] |= 2
                js = j + 1
            if j >= 1:
                pk[i], ps[i] = O.fe(ks[j - 1][0]), ks[j - 1][1]
                fl[i] |= 1
            if js < len(only):
                sk[i], ss[i] = O.fe(ks[js][0]), ks[js][1]
                fl[i] |= 4
        return pk, ps, sk, ss, fl

    def _rehash(self):
        self.tree = O.tree_build(O.hash3(self.pre, 2), 2)

    def shard_insert_apply(self, x, upd, local_depth):
        x = np.asarray(x, dtype=np.uint64)
        base = self.rank * self.n
        sub_roots = np.zeros((len(x), 4), np.uint64)
        sib = np.zeros((len(x), local_depth, 4), np.uint64)
        for t, g in enumerate(x):
            local = int(g) - base
            if 0 <= local < self.n:
                s, _ = O.get_proof(self.tree, self.n, local)       # the path of a leaf does not depend on the leaf itself
                sib[t] = s
                self.pre[local] = upd[t]
                self._rehash()                                     # the reference's own way: re-hash and rebuild (tiny trees only)
                sub_roots[t] = self.tree[-1]
        return sub_roots, sib

    def shard_insert_cap(self, x, sub_roots, cap_depth):
        x = np.asarray(x, dtype=np.uint64)
        leaves = np.array(O.levels(self.cap, self.world)[0]) if self.world > 1 else np.array([self.tree[-1]])
        roots = np.zeros((len(x), 4), np.uint64)
        sib = np.zeros((len(x), cap_depth, 4), np.uint64)
        for t, g in enumerate(x):
            owner = int(g) // self.n
            if self.world > 1:
                s, _ = O.get_proof(self.cap, self.world, owner)
                sib[t] = s
                leaves[owner] = sub_roots[t]
                self.cap = O.tree_build(np.ascontiguousarray(leaves), 1)
                roots[t] = self.cap[-1]
            else:
                roots[t] = sub_roots[t]
        return roots, sib


class OracleEngine:
    device = 0

    def shard_insert_plan(self, values, first_idx, pred_keys, pred_slots, succ_keys, succ_slots, flags):
        v = O.to_ints(values)
        b = len(v)
        pk = np.asarray(pred_keys, dtype=np.uint64).reshape(-1, b, 4)
        sk = np.asarray(succ_keys, dtype=np.uint64).reshape(-1, b, 4)
        ps, ss, fl = (np.asarray(a).reshape(-1, b) for a in (pred_slots, succ_slots, flags))
        x, upd = np.zeros(2 * b, np.uint64), np.zeros((2 * b, 3, 4), np.uint64)
        low_old, largest = np.zeros((b, 3, 4), np.uint64), np.zeros(b, np.uint8)
        for k in range(b):
            preds = [(O.to_int(pk[r, k]), int(ps[r, k])) for r in range(pk.shape[0]) if fl[r, k] & 1]
            succs = [(O.to_int(sk[r, k]), int(ss[r, k])) for r in range(pk.shape[0]) if fl[r, k] & 4]
            preds += [(v[i], first_idx + i) for i in range(k) if v[i] < v[k]]
            succs += [(v[i], first_idx + i) for i in range(k) if v[i] > v[k]]
            if v[k] == 0 or any(fl[r, k] & 2 for r in range(pk.shape[0])) or v[k] in v[:k] or not preds:
                raise ValueError("bad insert")
            p, s = max(preds), (min(succs) if succs else (0, 0))
            x[2 * k], x[2 * k + 1] = p[1], first_idx + k
            low_old[k] = O.fes([p[0], s[0], s[1]])
            upd[2 * k] = O.fes([p[0], v[k], first_idx + k])
            upd[2 * k + 1] = O.fes([v[k], s[0], s[1]])
            largest[k] = 0 if succs else 1
        return x, upd, low_old, largest

    def build_from_leaves(self, pre):
        return OracleTree(pre)

    def low_leaf_merge(self, values, cand_keys, cand_slots, flags, occupied_total, n_total, head_next_zero):
        v = O.to_ints(values)
        q = len(v)
        ck = np.asarray(cand_keys, dtype=np.uint64).reshape(-1, q, 4)
        cs, fl = np.asarray(cand_slots).reshape(-1, q), np.asarray(flags).reshape(-1, q)
        low, matched = np.zeros(q, np.uint64), np.zeros(q, bool)
        for i in range(q):
            cands = [(O.to_int(ck[r, i]), int(cs[r, i])) for r in range(ck.shape[0]) if fl[r, i] & 1]
            present = any(fl[r, i] & 2 for r in range(ck.shape[0]))
            if head_next_zero:
                matched[i] = True
            elif cands and not present:
                low[i], matched[i] = max(cands)[1], True
            elif v[i] != 0 and occupied_total < n_total:
                low[i], matched[i] = occupied_total, True
        return low, matched

    def trace_merkle_proofs(self, leaves, indices, siblings, want_states=True):
        roots = []
        for lf, ix, sb in zip(leaves, indices, siblings):
            h, ix = lf, int(ix)
            for s in sb:
                h = O.hash2(np.stack([h, s]) if ix % 2 == 0 else np.stack([s, h]))[0]
                ix //= 2
            roots.append(h)
        return np.stack(roots), None
