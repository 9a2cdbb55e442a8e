"""TEST INFRASTRUCTURE — pure-Python big-int restatement of the reference's hot path.

This is the *second* oracle (the first is oracle/imt_oracle.c).  It exists so the C oracle is
cross-checked by an implementation that shares no code and no number representation with it
(Python ints, no Montgomery form).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import anything under oracle/; the product (indexed-merkle-tree-halo2_b200/) never does.

Parity pin: Poseidon(0,0,0) literal at /root/reference/src/indexed_merkle_tree.rs:247-251 (the only
numeric known-answer the reference holds) — asserted in tests/test_oracle.py.  Per-round states are
"parity unpinned" by the reference (no reference test holds an intermediate state); they are pinned
here by naive-schedule == optimized-schedule on random states plus final-output equality.

The arithmetic is NOT in /root/reference: it lives in the un-vendored git dependencies
  pse-poseidon   (aerius-labs/pse-poseidon, branch feat/stateless-hash; Cargo.toml:16; commit unpinned)
  halo2curves    (BN254 Fr, named grumpkin::Fq at indexed_merkle_tree.rs:327)
so what follows restates the published Poseidon algorithm (Grain LFSR parameter generation, Cauchy MDS,
optimized round constants, sparse-MDS factorisation, sponge with capacity 2^64 and push-1 padding) and is
anchored on the reference's call sites:
  Poseidon::new(8,57)            indexed_merkle_tree.rs:370, 663, 681, 807
  update / squeeze_and_reset     utils.rs:46-47, 96-100; indexed_merkle_tree.rs:374-375, 667-668
  IndexedMerkleTree::new         utils.rs:20-57
  get_proof / verify_proof       utils.rs:63-85 / 87-107
  update_idx_leaf                indexed_merkle_tree.rs:632-660
  hash_nullifier_pre_images      indexed_merkle_tree.rs:662-671
  insert orchestration           indexed_merkle_tree.rs:710-741
"""

P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # indexed_merkle_tree.rs:383
T = 3
RATE = 2
R_F = 8
R_P = 57
KAT_H3_ZERO = 1960587138944869480785025106734196872454309951825657414575195034687326603497  # IMT:248


def inv(a):
    return pow(a % P, P - 2, P)


# --------------------------------------------------------------------------- Grain LFSR parameters
class Grain:
    """80-bit Grain LFSR in self-shrinking mode (Poseidon paper, appendix; pse-poseidon grain.rs)."""

    def __init__(self, n_bits, t, r_f, r_p):
        bits = []

        def append(width, value):
            for i in reversed(range(width)):
                bits.append((value >> i) & 1)

        append(2, 1)  # prime field
        append(4, 0)  # x^alpha S-box
        append(12, n_bits)
        append(12, t)
        append(10, r_f)
        append(10, r_p)
        append(30, (1 << 30) - 1)
        assert len(bits) == 80
        self.s = bits
        self.n_bits = n_bits
        for _ in range(160):
            self._new_bit()

    def _new_bit(self):
        s = self.s
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def next_bit(self):
        # pairs: keep the second bit only when the first is 1
        b = self._new_bit()
        while not b:
            self._new_bit()
            b = self._new_bit()
        return self._new_bit()

    def _next_int(self):
        v = 0
        for _ in range(self.n_bits):  # MSB first
            v = (v << 1) | self.next_bit()
        return v

    def next_fe(self):  # rejection sampling (round constants)
        while True:
            v = self._next_int()
            if v < P:
                return v

    def next_fe_no_reject(self):  # MDS x/y: reduced mod p
        return self._next_int() % P


def mat_mul(a, b):
    n = len(a)
    return [[sum(a[i][k] * b[k][j] for k in range(n)) % P for j in range(len(b[0]))] for i in range(n)]


def mat_vec(m, v):
    return [sum(m[i][j] * v[j] for j in range(len(v))) % P for i in range(len(m))]


def transpose(m):
    return [list(r) for r in zip(*m)]


def mat_inv(m):
    n = len(m)
    a = [list(r) + [1 if i == j else 0 for j in range(n)] for i, r in enumerate(m)]
    for c in range(n):
        piv = next(r for r in range(c, n) if a[r][c] % P)
        a[c], a[piv] = a[piv], a[c]
        k = inv(a[c][c])
        a[c] = [x * k % P for x in a[c]]
        for r in range(n):
            if r != c and a[r][c]:
                f = a[r][c]
                a[r] = [(x - f * y) % P for x, y in zip(a[r], a[c])]
    return [r[n:] for r in a]


class Spec:
    """Everything Poseidon::<Fr,3,2>::new(8,57) derives at construction (IMT:370)."""

    def __init__(self, r_f=R_F, r_p=R_P, t=T):
        g = Grain(254, t, r_f, r_p)
        self.r_f, self.r_p, self.t = r_f, r_p, t
        rc = [[g.next_fe() for _ in range(t)] for _ in range(r_f + r_p)]
        xs = [g.next_fe_no_reject() for _ in range(t)]
        ys = [g.next_fe_no_reject() for _ in range(t)]
        mds = [[inv(x + y) for y in ys] for x in xs]
        self.round_constants = rc
        self.mds = mds
        half = r_f // 2
        mds_inv = mat_inv(mds)
        # ---- optimized constants
        start = [rc[0]] + [mat_vec(mds_inv, rc[i]) for i in range(1, half)]
        partial = [0] * r_p
        acc = list(rc[half + r_p])
        for k in reversed(range(r_p)):
            tmp = mat_vec(mds_inv, acc)
            partial[k] = tmp[0]
            tmp[0] = 0
            acc = [(x + c) % P for x, c in zip(tmp, rc[half + k])]
        start.append(mat_vec(mds_inv, acc))
        end = [mat_vec(mds_inv, rc[i]) for i in range(half + r_p + 1, r_f + r_p)]
        self.start, self.partial, self.end = start, partial, end
        # ---- sparse factorisation of the partial-round linear layers
        mt = transpose(mds)
        acc_m = [list(r) for r in mt]
        sparse = []
        for _ in range(r_p):
            m_hat = [r[1:] for r in acc_m[1:]]
            w = [r[0] for r in acc_m[1:]]
            w_hat = mat_vec(mat_inv(m_hat), w)
            m_prime = [[1] + [0] * (t - 1)] + [[0] + list(r) for r in m_hat]
            pp = [list(acc_m[0])] + [[w_hat[i - 1]] + [1 if j == i else 0 for j in range(1, t)] for i in range(1, t)]
            ppt = transpose(pp)
            sparse.append((ppt[0], [ppt[i][0] for i in range(1, t)]))  # (row, col_hat)
            acc_m = mat_mul(mt, m_prime)
        sparse.reverse()
        self.sparse = sparse
        self.pre_sparse = transpose(acc_m)


_SPEC = None


def spec():
    global _SPEC
    if _SPEC is None:
        _SPEC = Spec()
    return _SPEC


# --------------------------------------------------------------------------- permutations
def sbox(x):
    x2 = x * x % P
    return x2 * x2 % P * x % P


def permute_naive(state, sp=None):
    """Textbook schedule: add round constants, S-box (full/partial), dense MDS. 828 modmuls."""
    sp = sp or spec()
    s = list(state)
    half = sp.r_f // 2
    for r in range(sp.r_f + sp.r_p):
        s = [(x + c) % P for x, c in zip(s, sp.round_constants[r])]
        if r < half or r >= half + sp.r_p:
            s = [sbox(x) for x in s]
        else:
            s[0] = sbox(s[0])
        s = mat_vec(sp.mds, s)
    return s


def permute_trace(state, sp=None, sbox_out=None):
    """Optimized schedule (pse-poseidon permutation.rs, same as halo2-base's in-circuit hasher).

    Returns (final_state, states) where states is the 66-entry witness trace of SURVEY §8a row 9:
    [after the pre-constant add] + [after the linear layer of each of the 4+57+4 rounds].
    sbox_out (optional list) receives the extended trace: (x^2, x^4, x^5 + c) of every S-box in execution order."""
    sp = sp or spec()
    half = sp.r_f // 2

    def sb(x, c):
        x2 = x * x % P
        x4 = x2 * x2 % P
        u = (x4 * x + c) % P
        if sbox_out is not None:
            sbox_out.append((x2, x4, u))
        return u

    s = [(x + c) % P for x, c in zip(state, sp.start[0])]
    out = [list(s)]
    for i in range(1, half):
        s = [sb(x, c) for x, c in zip(s, sp.start[i])]
        s = mat_vec(sp.mds, s)
        out.append(list(s))
    s = [sb(x, c) for x, c in zip(s, sp.start[half])]
    s = mat_vec(sp.pre_sparse, s)
    out.append(list(s))
    for k in range(sp.r_p):
        row, col_hat = sp.sparse[k]
        s0 = sb(s[0], sp.partial[k])
        v = [s0] + s[1:]
        s = [sum(r * x for r, x in zip(row, v)) % P] + [(col_hat[i - 1] * s0 + v[i]) % P for i in range(1, sp.t)]
        out.append(list(s))
    for i in range(half - 1):
        s = [sb(x, c) for x, c in zip(s, sp.end[i])]
        s = mat_vec(sp.mds, s)
        out.append(list(s))
    s = [sb(x, 0) for x in s]
    s = mat_vec(sp.mds, s)
    out.append(list(s))
    return s, out


def permute(state, sp=None):
    return permute_trace(state, sp)[0]


# --------------------------------------------------------------------------- sponge (pse-poseidon poseidon.rs)
class Poseidon:
    """Mirror of Poseidon::<Fr,T,RATE>: new / update / squeeze_and_reset (IMT:370-376). Default <3,2>(8,57);
    pass sp = Spec(r_f, r_p, t) for another instance (RATE = t - 1; utils.rs:6, 19 are generic over T and RATE)."""

    def __init__(self, r_f=R_F, r_p=R_P, sp=None):
        self.sp = sp or spec()
        assert sp is not None or (r_f, r_p) == (R_F, R_P)
        self.rate = self.sp.t - 1
        self._reset()

    def _reset(self):
        self.state = [1 << 64] + [0] * (self.sp.t - 1)
        self.absorbing = []

    def update(self, elements):
        buf = self.absorbing + [e % P for e in elements]
        self.absorbing = []
        for i in range(0, len(buf), self.rate):
            chunk = buf[i:i + self.rate]
            if len(chunk) < self.rate:
                self.absorbing = chunk
            else:
                for j, e in enumerate(chunk):
                    self.state[1 + j] = (self.state[1 + j] + e) % P
                self.state = permute(self.state, self.sp)

    def squeeze_and_reset(self):
        last = self.absorbing + [1]
        for j, e in enumerate(last):
            self.state[1 + j] = (self.state[1 + j] + e) % P
        self.state = permute(self.state, self.sp)
        out = self.state[1]
        self._reset()
        return out


def hash_n(inputs, sp=None):
    """update(inputs) + squeeze_and_reset() for any input length and instance"""
    h = Poseidon(sp=sp)
    h.update(list(inputs))
    return h.squeeze_and_reset()


def hash_trace_n(inputs, sp=None, sbox_out=None):
    """digest + every traced state ((len // rate + 1) x (1 + r_f + r_p) states of t values) of one hash, any instance;
    sbox_out (optional list) receives the extended S-box trace of all permutations"""
    sp = sp or spec()
    rate = sp.t - 1
    s = [1 << 64] + [0] * rate
    buf = [e % P for e in inputs]
    states = []
    while True:
        chunk, buf = buf[:rate], buf[rate:]
        last = len(chunk) < rate
        if last:
            chunk = chunk + [1]
        for j, e in enumerate(chunk):
            s[1 + j] = (s[1 + j] + e) % P
        s, tr = permute_trace(s, sp, sbox_out)
        states += tr
        if last:
            return s[1], states


def hash2(l, r):
    h = Poseidon()
    h.update([l, r])
    return h.squeeze_and_reset()


def hash3(a, b, c):
    h = Poseidon()
    h.update([a, b, c])
    return h.squeeze_and_reset()


def hash_trace(inputs):
    """The 132 states (2 permutations x 66) of one fixed-length hash of 2 or 3 inputs, plus the digest."""
    assert len(inputs) in (2, 3)
    s = [1 << 64, inputs[0] % P, inputs[1] % P]
    s, t1 = permute_trace(s)
    if len(inputs) == 3:
        s = [s[0], (s[1] + inputs[2]) % P, (s[2] + 1) % P]
    else:
        s = [s[0], (s[1] + 1) % P, s[2]]
    s, t2 = permute_trace(s)
    return s[1], t1 + t2


# --------------------------------------------------------------------------- native tree (utils.rs)
class IndexedMerkleTree:
    def __init__(self, leaves, h2=None):  # utils.rs:20-57; h2: node hash of another Poseidon instance
        hash2 = self._h2 = h2 or globals()["hash2"]
        if len(leaves) == 0:
            raise ValueError("Cannot create Merkle Tree with no leaves")
        if len(leaves) == 1:
            self.tree, self.root = [list(leaves)], leaves[0]
            return
        if len(leaves) % 2 == 1:
            raise ValueError("Leaves must be even")
        self.tree = [list(leaves)]
        cur = list(leaves)
        while len(cur) > 1:
            cur = [hash2(cur[i], cur[i + 1]) for i in range(0, len(cur), 2)]  # IndexError like UT:45
            self.tree.append(cur)
        self.root = cur[0]

    def get_root(self):  # utils.rs:59-61
        return self.root

    def get_proof(self, index):  # utils.rs:63-85
        proof, helper = [], []
        for lvl in self.tree[:-1]:
            left = index % 2 == 0
            proof.append(lvl[index + 1 if left else index - 1])
            helper.append(1 if left else 0)
            index //= 2
        return proof, helper

    def verify_proof(self, leaf, index, root, proof):  # utils.rs:87-107
        hash2 = self._h2
        h = leaf
        for sib in proof:
            h = hash2(h, sib) if index % 2 == 0 else hash2(sib, h)
            index //= 2
        return h == root


def hash_preimages(pre):  # indexed_merkle_tree.rs:662-671; order val, next_val, next_idx
    return [hash3(v, nv, ni) for (v, nv, ni) in pre]


def update_idx_leaf(leaves, new_val, new_val_idx):  # indexed_merkle_tree.rs:632-660
    out = [list(l) for l in leaves]
    low = 0
    for i, (val, next_val, _next_idx) in enumerate(leaves):
        if next_val == 0 and i == 0:
            out[i + 1][0] = new_val
            out[i][1] = new_val
            out[i][2] = i + 1
            low = i
            break
        if val < new_val and (next_val > new_val or next_val == 0):
            out[new_val_idx][0] = new_val
            out[new_val_idx][1] = out[i][1]
            out[new_val_idx][2] = out[i][2]
            out[i][1] = new_val
            out[i][2] = new_val_idx
            low = i
            break
    return out, low


def insert_rounds(depth, new_vals, start_idx=1):
    """indexed_merkle_tree.rs:679-803 (native half): returns per-round witness dicts."""
    n = 1 << depth
    pre = [[0, 0, 0] for _ in range(n)]
    tree = IndexedMerkleTree(hash_preimages(pre))
    rounds = []
    for r, v in enumerate(new_vals):
        idx = start_idx + r
        old_root = tree.get_root()
        new_pre, low = update_idx_leaf(pre, v, idx)
        low_leaf = list(pre[low])
        low_proof, low_helper = tree.get_proof(low)
        tree = IndexedMerkleTree(hash_preimages(new_pre))
        new_proof, new_helper = tree.get_proof(idx)
        rounds.append(dict(old_root=old_root, low_idx=low, low_leaf=low_leaf, low_proof=low_proof,
                           low_helper=low_helper, new_root=tree.get_root(), new_leaf=list(new_pre[idx]),
                           new_idx=idx, new_proof=new_proof, new_helper=new_helper,
                           is_largest=1 if new_pre[idx][1] == 0 else 0))
        pre = new_pre
    return rounds, pre
