/* TEST: a plain C99 client of include/imt_b200.h — what a cgo / Rust `extern "C"` binding sees: no CUDA headers, no
 * C++, only pointers and sizes. Built by tests/test_cabi_driver.py with gcc and linked against libimt_b200.so.
 *   exit 0  everything checked (GPU present)
 *   exit 3  imt_ctx_create failed cleanly (no GPU): the library has no CPU fallback
 * The scenario is the native half of the reference's test_insert_leaf_multiple_round
 * (/root/reference/src/indexed_merkle_tree.rs:679-741): empty depth-3 tree, inserts 30,10,20,5,50,35. */
#include <inttypes.h>
#include <stdio.h>
#include <string.h>

#include "imt_b200.h"

#define CHECK(call)                                                                              \
    do {                                                                                         \
        imt_status st_ = (call);                                                                 \
        if (st_ != IMT_OK) {                                                                     \
            fprintf(stderr, "%s -> %d (%s) %s\n", #call, (int)st_, imt_status_string(st_), imt_last_error(ctx)); \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

static void print_fe(const char* tag, const uint64_t* fe) {
    printf("%s %016" PRIx64 "%016" PRIx64 "%016" PRIx64 "%016" PRIx64 "\n", tag, fe[3], fe[2], fe[1], fe[0]);
}

int main(void) {
    imt_ctx* ctx = NULL;
    imt_status st = imt_ctx_create(0, IMT_FE_CANONICAL, &ctx);
    if (st != IMT_OK) {
        printf("no-gpu status=%d\n", (int)st);
        return ctx == NULL && st == IMT_ERR_CUDA ? 3 : 1;
    }
    /* test_hash_zero (IMT:805-810): Poseidon(0,0,0), the literal at IMT:247-251 */
    uint64_t zeros[12] = {0}, kat[4];
    CHECK(imt_poseidon_hash3(ctx, zeros, 1, kat));
    print_fe("kat", kat);

    /* error strings of IndexedMerkleTree::new (utils.rs:25, 35) */
    imt_tree* tree = NULL;
    uint64_t three[12] = {1, 0, 0, 0, 2, 0, 0, 0, 3, 0, 0, 0};
    st = imt_tree_build_from_hashes(ctx, three, 0, &tree);
    printf("empty: %d %s\n", (int)st, imt_status_string(st));
    st = imt_tree_build_from_hashes(ctx, three, 3, &tree);
    printf("odd: %d %s\n", (int)st, imt_status_string(st));

    /* 8 empty leaves, then six inserts in one batch with every witness the chip loads */
    uint64_t pre[8 * 12];
    memset(pre, 0, sizeof pre);
    CHECK(imt_tree_build_from_leaves(ctx, pre, 8, &tree));
    uint64_t root[4];
    CHECK(imt_tree_root(tree, root));
    print_fe("empty_root", root);
    size_t occupied = 0;
    CHECK(imt_tree_occupied(tree, &occupied));
    const uint64_t ins[6] = {30, 10, 20, 5, 50, 35};
    uint64_t vals[6 * 4] = {0};
    for (int i = 0; i < 6; ++i) vals[4 * i] = ins[i];
    uint64_t old_roots[6 * 4], new_roots[6 * 4], low_idx[6], low_leaves[6 * 12], new_leaves[6 * 12], low_sib[6 * 3 * 4], new_sib[6 * 3 * 4];
    uint8_t low_hel[6 * 3], new_hel[6 * 3], largest[6];
    static uint64_t fold_nodes[6 * 4 * 3 * 4]; /* chain values of the four folds: make the witness trace below one launch */
    imt_insert_witness w = {old_roots, low_idx, low_leaves, low_sib, low_hel, new_roots, new_leaves, new_sib, new_hel, largest, fold_nodes};
    CHECK(imt_insert_batch(tree, vals, 6, occupied, &w));
    for (int i = 0; i < 6; ++i) {
        printf("low_idx %" PRIu64 " largest %d\n", low_idx[i], (int)largest[i]);
        print_fe("root", new_roots + 4 * i);
    }
    /* verify_proof (utils.rs:87-107) of every new leaf under its new root, through the batched call */
    uint64_t leaf_hashes[6 * 4], idx[6];
    uint8_t ok[6];
    CHECK(imt_poseidon_hash3(ctx, new_leaves, 6, leaf_hashes));
    for (int i = 0; i < 6; ++i) idx[i] = occupied + (uint64_t)i;
    CHECK(imt_verify_proofs(ctx, leaf_hashes, idx, new_roots, new_sib, 6, 3, ok));
    for (int i = 0; i < 6; ++i) {
        if (!ok[i]) {
            fprintf(stderr, "new leaf %d does not verify\n", i);
            return 1;
        }
    }
    /* non-inclusion of 25: low leaf 20 -> 30 */
    uint64_t q[4] = {25, 0, 0, 0}, li, ll[12], sib[3 * 4];
    uint8_t matched, hel[3], lg;
    CHECK(imt_non_inclusion_paths(tree, q, 1, &li, &matched, ll, sib, hel, &lg));
    printf("non_inclusion low_idx %" PRIu64 " val %" PRIu64 " next %" PRIu64 " matched %d largest %d\n", li, ll[0], ll[4], (int)matched, (int)lg);
    /* verify_merkle_proof witness of leaf 3 straight from the tree: depth x 132 x 3 FE; the last state holds the root */
    static uint64_t states[3 * 132 * 3 * 4];
    uint64_t leaf3 = 3;
    CHECK(imt_tree_trace_proofs(tree, &leaf3, 1, states));
    CHECK(imt_tree_root(tree, root));
    printf("trace_ends_in_root %d\n", memcmp(states + ((2 * 132 + 131) * 3 + 1) * 4, root, 32) == 0);
    /* 128-bit limb witnesses of verify_non_inclusion for (low leaf 20 -> 30, new value 25) */
    uint64_t limbs[6 * 4];
    uint8_t flags[3];
    CHECK(imt_non_inclusion_limbs(ctx, ll, q, 1, limbs, flags));
    printf("limbs nl_r %" PRIu64 " ll_r %" PRIu64 " llv_r %" PRIu64 " flags %d%d%d\n", limbs[4], limbs[12], limbs[20], flags[0], flags[1], flags[2]);
    imt_tree_destroy(tree);
    /* another Poseidon instance (utils.rs:6, 19 are generic over T and RATE): <5, 4>(8, 60); the published permutation
     * vector poseidonperm_x5_254_5 of [0, 1, 2, 3, 4], and a 6-element update + squeeze_and_reset */
    imt_ctx* ctx5 = NULL;
    st = imt_ctx_create_spec(0, IMT_FE_CANONICAL, 5, 4, 8, 60, &ctx5);
    if (st != IMT_OK) {
        fprintf(stderr, "imt_ctx_create_spec -> %d\n", (int)st);
        return 1;
    }
    uint64_t in5[6 * 4] = {0}, out5[5 * 4], dg[4];
    for (int i = 0; i < 5; ++i) in5[4 * i] = (uint64_t)i;
    if (imt_poseidon_permute(ctx5, in5, 1, out5) != IMT_OK) return 1;
    print_fe("perm5", out5);
    for (int i = 0; i < 6; ++i) in5[4 * i] = (uint64_t)i + 1;
    if (imt_poseidon_hash(ctx5, in5, 6, 1, dg) != IMT_OK) return 1;
    print_fe("hash5_6", dg);
    size_t fe_per_hash = 0;
    if (imt_trace_fe_per_hash(ctx5, 6, &fe_per_hash) != IMT_OK) return 1;
    printf("trace_fe %zu\n", fe_per_hash);
    imt_ctx_destroy(ctx5);
    /* checkpoint: save, inspect the header without a device, load into a fresh tree, same root and preimages */
    CHECK(imt_tree_build_from_leaves(ctx, pre, 8, &tree));
    CHECK(imt_tree_occupied(tree, &occupied));
    CHECK(imt_insert_batch(tree, vals, 6, occupied, &w));
    const char* path = "/tmp/imt_b200_cabi_driver.ckpt";
    CHECK(imt_tree_save(tree, path));
    imt_checkpoint_info info;
    CHECK(imt_checkpoint_read_info(path, &info));
    imt_tree* back = NULL;
    CHECK(imt_tree_load(ctx, path, &back));
    uint64_t root2[4], pre_a[8 * 12], pre_b[8 * 12];
    CHECK(imt_tree_root(tree, root));
    CHECK(imt_tree_root(back, root2));
    CHECK(imt_tree_preimages(tree, pre_a));
    CHECK(imt_tree_preimages(back, pre_b));
    printf("checkpoint n %" PRIu64 " depth %u same_root %d same_leaves %d header_root %d\n", info.num_leaves, info.depth,
           memcmp(root, root2, 32) == 0, memcmp(pre_a, pre_b, sizeof pre_a) == 0, memcmp(info.root, root, 32) == 0);
    FILE* f = fopen(path, "r+b"); /* flip one bit of leaf 2: the load must refuse the file */
    if (!f) return 1;
    fseek(f, 64 + 96 * 2, SEEK_SET);
    int c = fgetc(f);
    fseek(f, 64 + 96 * 2, SEEK_SET);
    fputc(c ^ 1, f);
    fclose(f);
    imt_tree* corrupt = NULL;
    st = imt_tree_load(ctx, path, &corrupt);
    printf("corrupt checkpoint: %d %s\n", (int)st, imt_last_error(ctx));
    remove(path);
    imt_tree_destroy(back);
    /* the whole insert_leaf witness trace of the batch in one call: 3 + 4 x depth hashes per insert (IMT:253-313) */
    static uint64_t tr_states[6 * 15 * 132 * 3 * 4];
    uint64_t tr_roots[6 * 4 * 4], tr_newlow[6 * 12], tr_limbs[6 * 6 * 4];
    uint8_t tr_flags[6 * 3];
    CHECK(imt_insert_witness_trace(ctx, &w, 6, 3, occupied, tr_states, tr_roots, tr_newlow, tr_limbs, tr_flags));
    int trace_ok = imt_insert_trace_hashes(3) == 15;
    for (int i = 0; i < 6; ++i) {
        trace_ok &= memcmp(tr_roots + 16 * i, old_roots + 4 * i, 32) == 0;            /* fold of the low leaf        -> old root */
        trace_ok &= memcmp(tr_roots + 16 * i + 4, tr_roots + 16 * i + 8, 32) == 0;    /* both routes to the interim root */
        trace_ok &= memcmp(tr_roots + 16 * i + 12, new_roots + 4 * i, 32) == 0;       /* fold of the new leaf        -> new root */
        /* the last state of the last hash of the insert's trace holds the new root in element 1 */
        trace_ok &= memcmp(tr_states + ((((size_t)i * 15 + 14) * 132 + 131) * 3 + 1) * 4, new_roots + 4 * i, 32) == 0;
        trace_ok &= tr_flags[3 * i + 2] == 1;
    }
    /* the same trace without the chain values (1 + depth dependent launches) is identical */
    static uint64_t tr_states_loop[6 * 15 * 132 * 3 * 4];
    uint64_t tr_roots_loop[6 * 4 * 4];
    w.fold_nodes = NULL;
    CHECK(imt_insert_witness_trace(ctx, &w, 6, 3, occupied, tr_states_loop, tr_roots_loop, NULL, NULL, NULL));
    trace_ok &= memcmp(tr_states, tr_states_loop, sizeof tr_states) == 0 && memcmp(tr_roots, tr_roots_loop, sizeof tr_roots) == 0;
    printf("insert_trace_ok %d\n", trace_ok);
    /* the whole verify_non_inclusion witness of two values in one call (IMT:127-229): lookup, low leaf, path, limbs and the
     * Poseidon states of H3(low leaf) + its fold up the path = 1 + depth traced hashes per value */
    const uint64_t ni_vals[2 * 4] = {25, 0, 0, 0, 60, 0, 0, 0};
    uint64_t ni_low[2], ni_leaves[2 * 12], ni_sib[2 * 3 * 4], ni_limbs[2 * 6 * 4];
    uint8_t ni_matched[2], ni_largest[2], ni_flags[2 * 3];
    static uint64_t ni_states[2 * 4 * 132 * 3 * 4];
    CHECK(imt_non_inclusion_witness_trace(tree, ni_vals, 2, ni_low, ni_matched, ni_leaves, ni_sib, NULL, ni_largest, ni_limbs, ni_flags, ni_states));
    CHECK(imt_tree_root(tree, root));
    int ni_ok = imt_non_inclusion_trace_hashes(3) == 4 && ni_low[0] == li && memcmp(ni_leaves, ll, sizeof ll) == 0 && memcmp(ni_sib, sib, sizeof sib) == 0;
    for (int i = 0; i < 2; ++i) /* the last state of the last hash holds the root in element 1 */
        ni_ok &= memcmp(ni_states + ((((size_t)i * 4 + 3) * 132 + 131) * 3 + 1) * 4, root, 32) == 0 && ni_matched[i] == 1 && ni_flags[3 * i + 2] == 1;
    ni_ok &= ni_largest[0] == 0 && ni_largest[1] == 1 && ni_leaves[12] == 50; /* 60 goes behind the largest value, 50 */
    printf("non_inclusion_trace_ok %d\n", ni_ok);
    imt_tree_destroy(tree);
    imt_ctx_destroy(ctx);
    return 0;
}
