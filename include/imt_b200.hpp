// imt_b200.hpp — C++17 host-side mirror of the reference's Rust API for the hot path, header-only, above the C-ABI
// (imt_b200.h). No CUDA headers, no torch: links against libimt_b200.so only.
//
// The reference is a Rust crate and this image has no Rust toolchain, so the host side a Rust user would call is written
// here in C++ with the reference's names, argument meaning and error behaviour (the Rust `extern "C"` wrappers a
// maintainer adds are in INTEGRATION.md):
//   Poseidon<T, RATE>{new_(r_f, r_p), update, squeeze_and_reset}           pse-poseidon, call sites src/indexed_merkle_tree.rs:370-376
//   IndexedMerkleTreeLeaf{val, next_val, next_idx}                          src/utils.rs:12-17
//   IndexedMerkleTree<T, RATE>{new_, get_root, get_proof, verify_proof}     src/utils.rs:5-10, 20, 59, 63, 87
//   hash_nullifier_pre_images / update_idx_leaf                             src/indexed_merkle_tree.rs:662-671 / 632-660 (test helpers)
// plus the batched calls that replace the reference's per-element loops (insert_batch, non_inclusion_paths, traces).
// `new` is a C++ keyword: constructors that the reference calls `new` are `new_` here. A Rust `Result<T, &'static str>` is
// `Result<T>`; a Rust panic (index out of bounds, utils.rs:45, 76) is a thrown std::out_of_range. There is no CPU
// fallback: without a CUDA device `Poseidon::new_` throws NoDevice.
#pragma once
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "imt_b200.h"

namespace imt_b200 {

struct NoDevice : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct Error : std::runtime_error {
    imt_status status;
    Error(imt_status st, const std::string& what) : std::runtime_error(what), status(st) {}
};

// BN254 Fr as the canonical integer, 4 little-endian 64-bit words — halo2curves' `to_repr()` bytes (IMT_FE_CANONICAL).
struct Fr {
    std::array<uint64_t, 4> l{};
    static Fr zero() { return Fr{}; }
    static Fr one() { return from(1); }
    static Fr from(uint64_t v) {  // `Fr::from(30)` in the reference's tests
        Fr f;
        f.l[0] = v;
        return f;
    }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    friend bool operator==(const Fr& a, const Fr& b) { return a.l == b.l; }
    friend bool operator!=(const Fr& a, const Fr& b) { return !(a == b); }
    friend bool operator<(const Fr& a, const Fr& b) {  // `Fr: Ord` compares the canonical value (used at IMT:647)
        for (int i = 3; i >= 0; --i)
            if (a.l[i] != b.l[i]) return a.l[i] < b.l[i];
        return false;
    }
    friend bool operator>(const Fr& a, const Fr& b) { return b < a; }
    std::string hex() const {
        char buf[67];
        std::snprintf(buf, sizeof buf, "0x%016llx%016llx%016llx%016llx", (unsigned long long)l[3], (unsigned long long)l[2],
                      (unsigned long long)l[1], (unsigned long long)l[0]);
        return buf;
    }
};
static_assert(sizeof(Fr) == IMT_FE_BYTES, "Fr must be the 32-byte field element of the C-ABI");

// Rust's Result<T, &'static str>
template <class T>
class Result {
  public:
    static Result ok(T v) {
        Result r;
        r.v_ = std::move(v);
        return r;
    }
    static Result err(const char* e) {
        Result r;
        r.e_ = e;
        return r;
    }
    bool is_ok() const { return v_.has_value(); }
    bool is_err() const { return !is_ok(); }
    T unwrap() {
        if (!v_) throw std::runtime_error(std::string("called `Result::unwrap()` on an `Err` value: ") + e_);
        return std::move(*v_);
    }
    const char* unwrap_err() const { return e_; }

  private:
    std::optional<T> v_;
    const char* e_ = "";
};

namespace detail {
struct CtxDeleter {
    void operator()(imt_ctx* c) const { imt_ctx_destroy(c); }
};
inline void check(imt_ctx* ctx, imt_status st) {
    if (st == IMT_OK) return;
    const char* msg = imt_last_error(ctx);
    throw Error(st, (msg && *msg) ? msg : imt_status_string(st));
}
}  // namespace detail

// Poseidon::<Fr, T, RATE>::new(r_f, r_p): the hasher object owns the GPU context of that instance.
template <size_t T, size_t RATE>
class Poseidon {
  public:
    static Poseidon new_(size_t r_f, size_t r_p, int device = 0) { return Poseidon(r_f, r_p, device); }
    // buffers the elements; the sponge state is only ever materialised on the GPU (one kernel per squeeze)
    void update(const std::vector<Fr>& elements) { buf_.insert(buf_.end(), elements.begin(), elements.end()); }
    void update(const Fr* elements, size_t n) { buf_.insert(buf_.end(), elements, elements + n); }
    Fr squeeze_and_reset() {
        Fr out;
        std::vector<Fr> in;
        in.swap(buf_);
        detail::check(ctx(), imt_poseidon_hash(ctx(), in.data(), in.size(), 1, &out));
        return out;
    }
    // batched form of `update(&x[i*arity..]); squeeze_and_reset()` — one kernel for all n hashes
    std::vector<Fr> hash_many(const std::vector<Fr>& inputs, size_t arity) {
        const size_t n = arity ? inputs.size() / arity : 0;
        std::vector<Fr> out(n);
        detail::check(ctx(), imt_poseidon_hash(ctx(), inputs.data(), arity, n, out.data()));
        return out;
    }
    imt_ctx* ctx() const { return ctx_.get(); }
    size_t r_f() const { return r_f_; }
    size_t r_p() const { return r_p_; }

  private:
    Poseidon(size_t r_f, size_t r_p, int device) : r_f_(r_f), r_p_(r_p) {
        imt_ctx* c = nullptr;
        // <3, 2>(8, 57) is the instance the reference instantiates (IMT:362-365): tuned kernels; anything else: any-width
        const imt_status st = (T == 3 && RATE == 2 && r_f == 8 && r_p == 57)
                                  ? imt_ctx_create(device, IMT_FE_CANONICAL, &c)
                                  : imt_ctx_create_spec(device, IMT_FE_CANONICAL, (unsigned)T, (unsigned)RATE, (unsigned)r_f, (unsigned)r_p, &c);
        if (st == IMT_ERR_CUDA) throw NoDevice("imt_ctx_create: no usable CUDA device (the GPU library has no CPU fallback)");
        if (st != IMT_OK) throw Error(st, "unsupported Poseidon instance: need 2 <= T <= 5, RATE = T - 1, R_F even, R_F + R_P <= 256");
        ctx_.reset(c);
    }
    std::unique_ptr<imt_ctx, detail::CtxDeleter> ctx_;
    std::vector<Fr> buf_;
    size_t r_f_, r_p_;
};

// utils.rs:12-17 — field order val, next_val, next_idx (also the hash order, IMT:667)
struct IndexedMerkleTreeLeaf {
    Fr val, next_val, next_idx;
};
static_assert(sizeof(IndexedMerkleTreeLeaf) == 3 * IMT_FE_BYTES, "leaf = 3 FE, AoS");

// Everything the chip's insert_leaf loads with ctx.load_witness (IMT:444-489), per insert of a batch
struct InsertWitness {
    std::vector<Fr> old_roots, new_roots;
    std::vector<uint64_t> low_idx;
    std::vector<IndexedMerkleTreeLeaf> low_leaves, new_leaves;
    std::vector<std::vector<Fr>> low_proof, new_proof;          // siblings bottom-up
    std::vector<std::vector<Fr>> low_proof_helper, new_proof_helper;  // Fr::one() when the node is LEFT (utils.rs:70, 79)
    std::vector<bool> is_new_leaf_largest;                      // IMT:736-741
    std::vector<Fr> fold_nodes;  // [b][4][depth]: chain values of the four folds insert_leaf constrains (imt_b200.h); lets trace_insert_witness run as one launch
};

template <size_t T, size_t RATE>
class IndexedMerkleTree {
  public:
    // IndexedMerkleTree::new(hash, leaves) -> Result<Self, &'static str>            utils.rs:20-57
    static Result<IndexedMerkleTree> new_(Poseidon<T, RATE>& hash, std::vector<Fr> leaves) {
        imt_tree* t = nullptr;
        const imt_status st = imt_tree_build_from_hashes(hash.ctx(), leaves.data(), leaves.size(), &t);
        return wrap(hash, t, st, leaves.size());
    }
    // leaf hashing (IMT:662-671) fused in front of the build; keeps the preimages on the device, which the indexed calls
    // (low-leaf lookups, inserts) need
    static Result<IndexedMerkleTree> from_preimages(Poseidon<T, RATE>& hash, const std::vector<IndexedMerkleTreeLeaf>& leaves) {
        imt_tree* t = nullptr;
        const imt_status st = imt_tree_build_from_leaves(hash.ctx(), leaves.data(), leaves.size(), &t);
        return wrap(hash, t, st, leaves.size());
    }
    IndexedMerkleTree(IndexedMerkleTree&&) = default;
    IndexedMerkleTree& operator=(IndexedMerkleTree&&) = default;

    Fr get_root() const { return root_; }                                              // utils.rs:59-61
    size_t depth() const { return imt_tree_depth(tree_.get()); }
    // tree: Vec<Vec<F>> (utils.rs:8), one level at a time
    std::vector<Fr> level(unsigned lvl) const {
        std::vector<Fr> out(n_ >> lvl);
        detail::check(ctx_, imt_tree_level(tree_.get(), lvl, out.data()));
        return out;
    }
    // (siblings bottom-up, helper = 1 when the current node is the LEFT child)          utils.rs:63-85
    std::pair<std::vector<Fr>, std::vector<Fr>> get_proof(size_t index) const {
        if (index >= n_) throw std::out_of_range("index out of bounds");               // the reference panics at utils.rs:76
        const size_t d = depth();
        std::vector<Fr> sib(d), hel(d);
        const uint64_t idx = index;
        detail::check(ctx_, imt_tree_get_proofs_fe(tree_.get(), &idx, 1, sib.data(), hel.data()));
        return {std::move(sib), std::move(hel)};
    }
    bool verify_proof(const Fr& leaf, size_t index, const Fr& root, const std::vector<Fr>& proof) {   // utils.rs:87-107
        const uint64_t idx = index;
        uint8_t ok = 0;
        detail::check(ctx_, imt_verify_proofs(ctx_, &leaf, &idx, &root, proof.data(), 1, (unsigned)proof.size(), &ok));
        return ok == 1;
    }

    // ---- batched replacements of the reference's per-element loops (trees made by from_preimages)
    size_t occupied() const {
        size_t m = 0;
        detail::check(ctx_, imt_tree_occupied(tree_.get(), &m));
        return m;
    }
    // the read-only half of update_idx_leaf (IMT:632-660) for many values at once
    std::vector<uint64_t> low_leaf_lookup(const std::vector<Fr>& values) const {
        std::vector<uint64_t> low(values.size());
        std::vector<uint8_t> matched(values.size());
        detail::check(ctx_, imt_low_leaf_lookup(tree_.get(), values.data(), values.size(), low.data(), matched.data()));
        return low;
    }
    std::vector<IndexedMerkleTreeLeaf> preimages() const {
        std::vector<IndexedMerkleTreeLeaf> out(n_);
        detail::check(ctx_, imt_tree_preimages(tree_.get(), out.data()));
        return out;
    }
    // witnesses of verify_non_inclusion (IMT:127-137) for many values at once: low leaf, its path, is_new_leaf_largest,
    // and the 128-bit limb splits the chip assigns (IMT:143-172, 206-222) with the outcome of its two comparisons
    struct NonInclusion {
        std::vector<uint64_t> low_idx;
        std::vector<IndexedMerkleTreeLeaf> low_leaves;
        std::vector<std::vector<Fr>> low_proof, low_proof_helper;
        std::vector<bool> is_new_leaf_largest, valid;            // valid: passes the chip's prover-side assertions (IMT:190, 226-228)
        std::vector<std::array<Fr, 6>> limbs;                    // nl_q, nl_r, ll_q, ll_r, llv_q, llv_r
    };
    NonInclusion non_inclusion_paths(const std::vector<Fr>& values) const {
        const size_t q = values.size(), d = depth();
        NonInclusion o;
        o.low_idx.resize(q), o.low_leaves.resize(q), o.limbs.resize(q);
        std::vector<uint8_t> matched(q), hel(q * d), lg(q), flags(3 * q);
        std::vector<Fr> sib(q * d);
        detail::check(ctx_, imt_non_inclusion_paths(tree_.get(), values.data(), q, o.low_idx.data(), matched.data(), o.low_leaves.data(), sib.data(),
                                                    hel.data(), lg.data()));
        detail::check(ctx_, imt_non_inclusion_limbs(ctx_, o.low_leaves.data(), values.data(), q, o.limbs.data(), flags.data()));
        for (size_t i = 0; i < q; ++i) {
            o.low_proof.emplace_back(sib.begin() + i * d, sib.begin() + (i + 1) * d);
            std::vector<Fr> h(d);
            for (size_t k = 0; k < d; ++k) h[k] = Fr::from(hel[i * d + k]);
            o.low_proof_helper.push_back(std::move(h));
            o.is_new_leaf_largest.push_back(lg[i] != 0);
            o.valid.push_back(matched[i] != 0 && flags[3 * i + 2] != 0);
        }
        return o;
    }
    // The whole witness of verify_non_inclusion (IMT:127-229) in ONE call: the values above + the Poseidon states of the 1 + depth
    // hashes it constrains per value — [0] H3(low leaf) IMT:193-194, [1 .. depth] its fold up the path IMT:196-204 — flattened
    // [q][1 + depth][states per hash][T] Fr into `states`.
    NonInclusion non_inclusion_witness_trace(const std::vector<Fr>& values, std::vector<Fr>* states) const {
        const size_t q = values.size(), d = depth();
        NonInclusion o;
        o.low_idx.resize(q), o.low_leaves.resize(q), o.limbs.resize(q);
        std::vector<uint8_t> matched(q), hel(q * d), lg(q), flags(3 * q);
        std::vector<Fr> sib(q * d);
        size_t fe = 0;
        detail::check(ctx_, imt_trace_fe_per_hash(ctx_, 2, &fe));
        if (states) states->resize(q * imt_non_inclusion_trace_hashes((unsigned)d) * fe);
        detail::check(ctx_, imt_non_inclusion_witness_trace(tree_.get(), values.data(), q, o.low_idx.data(), matched.data(), o.low_leaves.data(),
                                                            sib.data(), hel.data(), lg.data(), o.limbs.data(), flags.data(),
                                                            states ? states->data() : nullptr));
        for (size_t i = 0; i < q; ++i) {
            o.low_proof.emplace_back(sib.begin() + i * d, sib.begin() + (i + 1) * d);
            std::vector<Fr> h(d);
            for (size_t k = 0; k < d; ++k) h[k] = Fr::from(hel[i * d + k]);
            o.low_proof_helper.push_back(std::move(h));
            o.is_new_leaf_largest.push_back(lg[i] != 0);
            o.valid.push_back(matched[i] != 0 && flags[3 * i + 2] != 0);
        }
        return o;
    }
    // IMT:710-741 for a whole batch with O(depth) hashes per insert: the tree advances in place and every per-insert
    // witness comes back — bit-identical to re-hashing and rebuilding per insert as the reference does (IMT:724-730)
    InsertWitness insert_batch(const std::vector<Fr>& new_vals) {
        const size_t b = new_vals.size(), d = depth();
        InsertWitness w;
        w.old_roots.resize(b), w.new_roots.resize(b), w.low_idx.resize(b), w.low_leaves.resize(b), w.new_leaves.resize(b);
        std::vector<Fr> ls(b * d), ns(b * d);
        std::vector<uint8_t> lh(b * d), nh(b * d), lg(b);
        w.fold_nodes.resize(b * 4 * d);
        imt_insert_witness cw{w.old_roots.data(), w.low_idx.data(), w.low_leaves.data(), ls.data(), lh.data(),
                              w.new_roots.data(), w.new_leaves.data(), ns.data(), nh.data(), lg.data(), w.fold_nodes.data()};
        detail::check(ctx_, imt_insert_batch(tree_.get(), new_vals.data(), b, occupied(), &cw));
        for (size_t i = 0; i < b; ++i) {
            w.low_proof.emplace_back(ls.begin() + i * d, ls.begin() + (i + 1) * d);
            w.new_proof.emplace_back(ns.begin() + i * d, ns.begin() + (i + 1) * d);
            std::vector<Fr> a(d), c(d);
            for (size_t k = 0; k < d; ++k) a[k] = Fr::from(lh[i * d + k]), c[k] = Fr::from(nh[i * d + k]);
            w.low_proof_helper.push_back(std::move(a)), w.new_proof_helper.push_back(std::move(c));
            w.is_new_leaf_largest.push_back(lg[i] != 0);
        }
        if (b) root_ = w.new_roots.back();
        return w;
    }
    // The Poseidon witness of everything insert_leaf hashes for a batch (IMT:253-313), one call: [b][3 + 4 depth][states per
    // hash][T] Fr flattened, in the chip's call order; `roots` (optional) receives old / interim / interim / new per insert.
    // `first_idx` = the slot of the batch's first insert (occupied() before insert_batch).
    std::vector<Fr> trace_insert_witness(const InsertWitness& w, size_t first_idx, std::vector<std::array<Fr, 4>>* roots = nullptr) const {
        const size_t b = w.low_idx.size(), d = depth();
        size_t fe = 0;
        detail::check(ctx_, imt_trace_fe_per_hash(ctx_, 2, &fe));
        std::vector<Fr> ls(b * d), ns(b * d), states(b * imt_insert_trace_hashes((unsigned)d) * fe);
        for (size_t i = 0; i < b; ++i)
            for (size_t k = 0; k < d; ++k) ls[i * d + k] = w.low_proof[i][k], ns[i * d + k] = w.new_proof[i][k];
        imt_insert_witness cw{};
        cw.low_idx = const_cast<uint64_t*>(w.low_idx.data());
        cw.low_leaves = const_cast<IndexedMerkleTreeLeaf*>(w.low_leaves.data());
        cw.new_leaves = const_cast<IndexedMerkleTreeLeaf*>(w.new_leaves.data());
        cw.low_siblings = ls.data();
        cw.new_siblings = ns.data();
        if (w.fold_nodes.size() == b * 4 * d && d) cw.fold_nodes = const_cast<Fr*>(w.fold_nodes.data());
        if (roots) roots->resize(b);
        detail::check(ctx_, imt_insert_witness_trace(ctx_, &cw, b, (unsigned)d, first_idx, states.data(), roots ? roots->data() : nullptr, nullptr,
                                                     nullptr, nullptr));
        return states;
    }
    // checkpoint: the leaves as the reference's serde derive orders them (utils.rs:12-17), canonical 32-byte LE, + root
    void save(const std::string& path) const { detail::check(ctx_, imt_tree_save(tree_.get(), path.c_str())); }
    static Result<IndexedMerkleTree> load(Poseidon<T, RATE>& hash, const std::string& path) {
        imt_tree* t = nullptr;
        const imt_status st = imt_tree_load(hash.ctx(), path.c_str(), &t);
        detail::check(hash.ctx(), st);
        return wrap(hash, t, st, imt_tree_num_leaves(t));
    }
    // witness of verify_merkle_proof (IMT:65-96) for leaves of this tree: [q][depth][states per hash][T] Fr, flattened
    std::vector<Fr> trace_proofs(const std::vector<uint64_t>& indices) const {
        size_t fe = 0;
        detail::check(ctx_, imt_trace_fe_per_hash(ctx_, 2, &fe));
        std::vector<Fr> states(indices.size() * depth() * fe);
        detail::check(ctx_, imt_tree_trace_proofs(tree_.get(), indices.data(), indices.size(), states.data()));
        return states;
    }

  private:
    struct TreeDeleter {
        void operator()(imt_tree* t) const { imt_tree_destroy(t); }
    };
    static Result<IndexedMerkleTree> wrap(Poseidon<T, RATE>& hash, imt_tree* t, imt_status st, size_t n) {
        if (st == IMT_ERR_EMPTY) return Result<IndexedMerkleTree>::err("Cannot create Merkle Tree with no leaves");   // utils.rs:25
        if (st == IMT_ERR_ODD) return Result<IndexedMerkleTree>::err("Leaves must be even");                           // utils.rs:35
        if (st == IMT_ERR_NOT_POW2) throw std::out_of_range("index out of bounds");   // even, not a power of two: the reference panics at utils.rs:45
        detail::check(hash.ctx(), st);
        IndexedMerkleTree tr;
        tr.ctx_ = hash.ctx();
        tr.tree_.reset(t);
        tr.n_ = n;
        detail::check(tr.ctx_, imt_tree_root(t, &tr.root_));
        return Result<IndexedMerkleTree>::ok(std::move(tr));
    }
    IndexedMerkleTree() = default;
    imt_ctx* ctx_ = nullptr;   // borrowed from the hasher, which must outlive the tree (the reference's lifetime 'a, utils.rs:5-7)
    std::unique_ptr<imt_tree, TreeDeleter> tree_;
    size_t n_ = 0;
    Fr root_;
};

// IndexedMerkleTree::new over ALL the GPUs of the box in one call (include/imt_b200.h, "multi-GPU inside the library"): one
// process, one thread; the tree is sharded by subtree, the subtree roots cross NVLink in one ncclAllGather issued by the
// library, every query addresses the global tree. <3, 2>(8, 57) only (the instance the reference instantiates, IMT:362-365).
class MultiGpu {
  public:
    explicit MultiGpu(const std::vector<int>& devices) {
        imt_multi* m = nullptr;
        if (imt_multi_create(devices.data(), (unsigned)devices.size(), IMT_FE_CANONICAL, &m) != IMT_OK)
            throw NoDevice("imt_multi_create failed: no CUDA devices (there is no CPU fallback), not a power of two of them, or no NCCL");
        m_.reset(m);
    }
    imt_multi* get() const { return m_.get(); }
    unsigned size() const { return imt_multi_size(m_.get()); }

  private:
    struct Deleter {
        void operator()(imt_multi* m) const { imt_multi_destroy(m); }
    };
    std::unique_ptr<imt_multi, Deleter> m_;
};

class ShardedIndexedMerkleTree {
  public:
    static Result<ShardedIndexedMerkleTree> from_preimages(MultiGpu& gpus, const std::vector<IndexedMerkleTreeLeaf>& leaves) {
        imt_mtree* t = nullptr;
        const imt_status st = imt_multi_build_from_leaves(gpus.get(), leaves.data(), leaves.size(), &t);
        if (st == IMT_ERR_EMPTY) return Result<ShardedIndexedMerkleTree>::err("Cannot create Merkle Tree with no leaves");   // utils.rs:25
        if (st == IMT_ERR_ODD) return Result<ShardedIndexedMerkleTree>::err("Leaves must be even");                           // utils.rs:35
        if (st != IMT_OK) throw Error(st, imt_multi_last_error(gpus.get()));
        ShardedIndexedMerkleTree tr;
        tr.gpus_ = gpus.get();
        tr.tree_.reset(t);
        tr.check(imt_mtree_root(t, &tr.root_));
        return Result<ShardedIndexedMerkleTree>::ok(std::move(tr));
    }
    ShardedIndexedMerkleTree(ShardedIndexedMerkleTree&&) = default;
    ShardedIndexedMerkleTree& operator=(ShardedIndexedMerkleTree&&) = default;
    Fr get_root() const { return root_; }                                              // utils.rs:59-61
    size_t depth() const { return imt_mtree_depth(tree_.get()); }
    std::pair<std::vector<Fr>, std::vector<Fr>> get_proof(size_t index) const {        // utils.rs:63-85
        if (index >= imt_mtree_num_leaves(tree_.get())) throw std::out_of_range("index out of bounds");
        const size_t d = depth();
        std::vector<Fr> sib(d), hel(d);
        std::vector<uint8_t> h8(d);
        const uint64_t idx = index;
        check(imt_mtree_get_proofs(tree_.get(), &idx, 1, sib.data(), h8.data()));
        for (size_t k = 0; k < d; ++k) hel[k] = Fr::from(h8[k]);
        return {std::move(sib), std::move(hel)};
    }
    std::vector<uint64_t> low_leaf_lookup(const std::vector<Fr>& values) const {
        std::vector<uint64_t> low(values.size());
        std::vector<uint8_t> matched(values.size());
        check(imt_mtree_low_leaf_lookup(tree_.get(), values.data(), values.size(), low.data(), matched.data()));
        return low;
    }
    void save(const std::string& path) const { check(imt_mtree_save(tree_.get(), path.c_str())); }

  private:
    struct Deleter {
        void operator()(imt_mtree* t) const { imt_mtree_destroy(t); }
    };
    void check(imt_status st) const {
        if (st != IMT_OK) throw Error(st, imt_multi_last_error(gpus_));
    }
    ShardedIndexedMerkleTree() = default;
    imt_multi* gpus_ = nullptr;
    std::unique_ptr<imt_mtree, Deleter> tree_;
    Fr root_;
};

// indexed_merkle_tree.rs:662-671, batched: one kernel for all leaves
template <size_t T, size_t RATE>
inline std::vector<Fr> hash_nullifier_pre_images(Poseidon<T, RATE>& hash, const std::vector<IndexedMerkleTreeLeaf>& leaves) {
    std::vector<Fr> out(leaves.size());
    detail::check(hash.ctx(), imt_poseidon_hash(hash.ctx(), leaves.data(), 3, leaves.size(), out.data()));
    return out;
}

// indexed_merkle_tree.rs:632-660, literally (host scan + rewiring): returns (updated leaves, low leaf index)
inline std::pair<std::vector<IndexedMerkleTreeLeaf>, size_t> update_idx_leaf(std::vector<IndexedMerkleTreeLeaf> leaves, const Fr& new_val,
                                                                              uint64_t new_val_idx) {
    std::vector<IndexedMerkleTreeLeaf> out = leaves;
    size_t low = 0;
    for (size_t i = 0; i < leaves.size(); ++i) {
        const IndexedMerkleTreeLeaf& l = leaves[i];
        if (l.next_val.is_zero() && i == 0) {                                          // IMT:640-646: the very first insert
            out[i + 1].val = new_val;
            out[i].next_val = new_val;
            out[i].next_idx = Fr::from(i + 1);
            low = i;
            break;
        }
        if (l.val < new_val && (l.next_val > new_val || l.next_val.is_zero())) {       // IMT:647
            out[new_val_idx].val = new_val;
            out[new_val_idx].next_val = out[i].next_val;
            out[new_val_idx].next_idx = out[i].next_idx;
            out[i].next_val = new_val;
            out[i].next_idx = Fr::from(new_val_idx);
            low = i;
            break;
        }
    }
    return {std::move(out), low};
}

}  // namespace imt_b200
