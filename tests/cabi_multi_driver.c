/* TEST: the multi-GPU half of include/imt_b200.h from plain C99 — no Python, no torch, no CUDA headers: what a Rust host
 * links. Builds the depth-D tree over the library's synthetic leaves SHARDED over N ranks and checks it against the same
 * tree built on one GPU in the same process: root, paths, low-leaf lookups, a batch of inserts.
 *
 *   cabi_multi_driver multi N [D]            ONE process drives N devices (imt_multi_create). Device i = i mod device count,
 *                                            so on a one-GPU box the group runs on the copy transport instead of NCCL.
 *   cabi_multi_driver rank R N IDFILE [D]    one process per GPU (imt_comm_create): rank 0 writes the NCCL id to IDFILE, the
 *                                            others wait for it. Needs N distinct GPUs (NCCL refuses to share one).
 * exit 0 = everything matched, 3 = no GPU (clean failure, no CPU fallback), 1 = mismatch / error.
 * The leaves are synth.random_preimages(2^D) of the Python package (splitmix64 stream, seed "IMT"), restated here, so the
 * printed root is comparable with tests/golden/golden.json build_roots[D].random. */
#define _POSIX_C_SOURCE 200809L
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "imt_b200.h"

static const uint64_t P_WORDS[4] = {0x43E1F593F0000001ull, 0x2833E84879B97091ull, 0xB85045B68181585Dull, 0x30644E72E131A029ull};

static uint64_t splitmix(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* field element e of the stream: four words, top masked to 62 bits, minus p if >= p (synth.field_elements) */
static void synth_fe(uint64_t seed, uint64_t e, uint64_t* out) {
    for (int k = 0; k < 4; ++k) out[k] = splitmix(seed, 4 * e + (uint64_t)k);
    out[3] &= 0x3FFFFFFFFFFFFFFFull;
    int ge = 1;
    for (int k = 3; k >= 0; --k) {
        if (out[k] != P_WORDS[k]) {
            ge = out[k] > P_WORDS[k];
            break;
        }
    }
    if (ge) {
        uint64_t borrow = 0;
        for (int k = 0; k < 4; ++k) {
            const uint64_t a = out[k], t = a - P_WORDS[k], t2 = t - borrow;
            borrow = (a < P_WORDS[k]) | (t < borrow);
            out[k] = t2;
        }
    }
}
static void print_fe(const char* tag, const uint64_t* fe) {
    printf("%s %016" PRIx64 "%016" PRIx64 "%016" PRIx64 "%016" PRIx64 "\n", tag, fe[3], fe[2], fe[1], fe[0]);
}
#define OK(call)                                                                                           \
    do {                                                                                                   \
        imt_status st_ = (call);                                                                           \
        if (st_ != IMT_OK) {                                                                               \
            fprintf(stderr, "%s -> %d (%s) %s\n", #call, (int)st_, imt_status_string(st_), errtext());     \
            return 1;                                                                                      \
        }                                                                                                  \
    } while (0)

static imt_ctx* g_ctx = NULL;
static imt_multi* g_multi = NULL;
static const char* errtext(void) { return g_multi ? imt_multi_last_error(g_multi) : (g_ctx ? imt_last_error(g_ctx) : ""); }

/* a small well-formed indexed tree for the lookup / insert checks: slot 0 = head, values 10, 20, ..., in slot order */
static void indexed_leaves(uint64_t* pre, size_t n, size_t occupied) {
    memset(pre, 0, n * 96);
    for (size_t i = 0; i < occupied; ++i) {
        pre[12 * i] = 10 * i;                                         /* val */
        pre[12 * i + 4] = (i + 1 < occupied) ? 10 * (i + 1) : 0;      /* next_val */
        pre[12 * i + 8] = (i + 1 < occupied) ? i + 1 : 0;             /* next_idx */
    }
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s multi N [depth] | rank R N IDFILE [depth]\n", argv[0]);
        return 2;
    }
    const int multi_mode = strcmp(argv[1], "multi") == 0;
    const unsigned world = (unsigned)atoi(multi_mode ? argv[2] : argv[3]);
    const unsigned rank = multi_mode ? 0 : (unsigned)atoi(argv[2]);
    const char* idfile = multi_mode ? NULL : argv[4];
    const unsigned depth = (unsigned)atoi((multi_mode ? (argc > 3 ? argv[3] : "16") : (argc > 5 ? argv[5] : "16")));
    const size_t n = (size_t)1 << depth, n_local = n / world;
    const uint64_t seed = 0x494D54;

    /* the reference tree: the whole thing on one GPU (rank 0 / the single process only) */
    imt_ctx* ref_ctx = NULL;
    imt_status st = imt_ctx_create(0, IMT_FE_CANONICAL, &ref_ctx);
    if (st != IMT_OK) {
        printf("no-gpu status=%d\n", (int)st);
        return st == IMT_ERR_CUDA ? 3 : 1;
    }
    g_ctx = ref_ctx;
    uint64_t* pre = (uint64_t*)malloc(n * 96);
    uint64_t* ipre = (uint64_t*)malloc(n * 96);
    if (!pre || !ipre) return 1;
    for (size_t i = 0; i < 3 * n; ++i) synth_fe(seed, i, pre + 4 * i);
    const size_t occupied = n / 2;
    indexed_leaves(ipre, n, occupied);
    imt_tree *ref = NULL, *iref = NULL;
    uint64_t ref_root[4], iref_root[4];
    if (multi_mode || rank == 0) {
        OK(imt_tree_build_from_leaves(ref_ctx, pre, n, &ref));
        OK(imt_tree_root(ref, ref_root));
        OK(imt_tree_build_from_leaves(ref_ctx, ipre, n, &iref));
        OK(imt_tree_root(iref, iref_root));
    }
    /* queries: a few leaf indices across all shards; lookup values between and equal to keys; 5 inserts */
    enum { Q = 8, B = 5 };
    uint64_t idx[Q], qv[Q * 4] = {0}, ins[B * 4] = {0};
    for (int i = 0; i < Q; ++i) idx[i] = ((uint64_t)i * (n / Q) + (uint64_t)(i * 7) % (n / Q)) % n;
    idx[Q - 1] = n - 1;
    const uint64_t lookups[Q] = {5, 15, 10, 0, 10 * (occupied - 1) + 3, 10 * (occupied / 2) + 1, 1, 10 * (occupied - 1)};
    for (int i = 0; i < Q; ++i) qv[4 * i] = lookups[i];
    const uint64_t newv[B] = {7, 10 * (occupied - 1) + 9, 3, 10 * (occupied / 2) + 5, 8};
    for (int i = 0; i < B; ++i) ins[4 * i] = newv[i];
    uint64_t* sib_ref = (uint64_t*)malloc((size_t)Q * depth * 32);
    uint64_t* sib_got = (uint64_t*)malloc((size_t)Q * depth * 32);
    uint8_t* hel_ref = (uint8_t*)malloc((size_t)Q * depth);
    uint8_t* hel_got = (uint8_t*)malloc((size_t)Q * depth);
    uint64_t low_ref[Q], low_got[Q], root_got[4], occ_got = 0;
    uint8_t m_ref[Q], m_got[Q];
    uint64_t nr_ref[B * 4], nr_got[B * 4], li_ref[B], li_got[B];
    uint64_t* ls_ref = (uint64_t*)malloc((size_t)B * depth * 32);
    uint64_t* ls_got = (uint64_t*)malloc((size_t)B * depth * 32);
    imt_insert_witness w_ref = {NULL, li_ref, NULL, ls_ref, NULL, nr_ref, NULL, NULL, NULL, NULL, NULL};
    imt_insert_witness w_got = {NULL, li_got, NULL, ls_got, NULL, nr_got, NULL, NULL, NULL, NULL, NULL};
    if (ref) {
        OK(imt_tree_get_proofs(ref, idx, Q, sib_ref, hel_ref));
        OK(imt_low_leaf_lookup(iref, qv, Q, low_ref, m_ref));
        OK(imt_insert_batch(iref, ins, B, occupied, &w_ref));
    }

    int bad = 0;
    if (multi_mode) {
        /* ---- one process, N devices */
        int devs[64];
        int count = 1;
        { /* device count without CUDA headers: imt_ctx_create fails beyond the last device */
            imt_ctx* probe = NULL;
            while (count < 64 && imt_ctx_create(count, IMT_FE_CANONICAL, &probe) == IMT_OK) imt_ctx_destroy(probe), ++count;
        }
        for (unsigned i = 0; i < world; ++i) devs[i] = (int)(i % (unsigned)count);
        OK(imt_multi_create(devs, world, IMT_FE_CANONICAL, &g_multi));
        printf("multi world %u devices %d transport %s\n", world, count, imt_multi_uses_nccl(g_multi) ? "nccl" : "copy");
        imt_mtree* mt = NULL;
        OK(imt_multi_build_from_leaves(g_multi, pre, n, &mt));
        OK(imt_mtree_root(mt, root_got));
        OK(imt_mtree_get_proofs(mt, idx, Q, sib_got, hel_got));
        imt_mtree* imt = NULL;
        OK(imt_multi_build_from_leaves(g_multi, ipre, n, &imt));
        OK(imt_mtree_low_leaf_lookup(imt, qv, Q, low_got, m_got));
        OK(imt_mtree_occupied(imt, &occ_got));
        OK(imt_mtree_insert_batch(imt, ins, B, occ_got, &w_got));
        uint64_t r2[4];
        OK(imt_mtree_rebuild_from_leaves(mt, pre)); /* the steady-state call */
        OK(imt_mtree_root(mt, r2));
        bad |= memcmp(r2, root_got, 32) != 0;
        imt_mtree_destroy(imt);
        imt_mtree_destroy(mt);
        imt_multi_destroy(g_multi);
        g_multi = NULL;
    } else {
        /* ---- one process per GPU */
        uint8_t id[IMT_COMM_ID_BYTES];
        if (rank == 0) {
            OK(imt_comm_unique_id(id));
            char tmp[1024];
            snprintf(tmp, sizeof tmp, "%s.tmp", idfile);
            FILE* f = fopen(tmp, "wb");
            if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) return 1;
            fclose(f);
            if (rename(tmp, idfile) != 0) return 1;
        } else {
            FILE* f = NULL;
            for (int tries = 0; tries < 600 && !(f = fopen(idfile, "rb")); ++tries) {
                struct timespec ts = {0, 100 * 1000 * 1000};
                nanosleep(&ts, NULL);
            }
            if (!f || fread(id, 1, sizeof id, f) != sizeof id) {
                fprintf(stderr, "rank %u: no id file\n", rank);
                return 1;
            }
            fclose(f);
        }
        imt_ctx* ctx = NULL;
        OK(imt_ctx_create((int)rank, IMT_FE_CANONICAL, &ctx));
        g_ctx = ctx;
        OK(imt_comm_create(ctx, rank, world, id));
        int ver = 0;
        OK(imt_comm_info(ctx, NULL, NULL, &ver));
        printf("rank %u of %u nccl %d\n", rank, world, ver);
        imt_tree *t = NULL, *it = NULL;
        OK(imt_sharded_build_from_leaves(ctx, pre + 12 * (size_t)rank * n_local, n_local, &t));
        OK(imt_tree_root(t, root_got));
        OK(imt_sharded_get_proofs(t, idx, Q, sib_got, hel_got));
        OK(imt_sharded_build_from_leaves(ctx, ipre + 12 * (size_t)rank * n_local, n_local, &it));
        OK(imt_sharded_low_leaf_lookup(it, qv, Q, low_got, m_got));
        OK(imt_sharded_occupied(it, &occ_got));
        OK(imt_sharded_insert_batch(it, ins, B, occ_got, &w_got));
        imt_tree_destroy(it);
        imt_tree_destroy(t);
        OK(imt_comm_destroy(ctx));
        imt_ctx_destroy(ctx);
        g_ctx = ref_ctx;
    }
    print_fe("root", root_got);
    printf("occupied %" PRIu64 "\n", occ_got);
    bad |= occ_got != occupied;
    if (ref) {
        bad |= memcmp(root_got, ref_root, 32) != 0;
        bad |= memcmp(sib_got, sib_ref, (size_t)Q * depth * 32) != 0 || memcmp(hel_got, hel_ref, (size_t)Q * depth) != 0;
        bad |= memcmp(low_got, low_ref, sizeof low_ref) != 0 || memcmp(m_got, m_ref, sizeof m_ref) != 0;
        bad |= memcmp(li_got, li_ref, sizeof li_ref) != 0 || memcmp(nr_got, nr_ref, sizeof nr_ref) != 0;
        bad |= memcmp(ls_got, ls_ref, (size_t)B * depth * 32) != 0;
        printf("sharded == single-GPU: root %d paths %d lookups %d inserts %d\n", memcmp(root_got, ref_root, 32) == 0,
               memcmp(sib_got, sib_ref, (size_t)Q * depth * 32) == 0, memcmp(low_got, low_ref, sizeof low_ref) == 0,
               memcmp(nr_got, nr_ref, sizeof nr_ref) == 0 && memcmp(ls_got, ls_ref, (size_t)B * depth * 32) == 0);
        imt_tree_destroy(ref);
        imt_tree_destroy(iref);
    }
    imt_ctx_destroy(ref_ctx);
    return bad ? 1 : 0;
}
