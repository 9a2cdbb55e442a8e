// Checkpoint of a built tree in the C-ABI: the leaves as a flat array of the reference's native struct
// `IndexedMerkleTreeLeaf { val, next_val, next_idx }` (/root/reference/src/utils.rs:12-17, serde-derived) — three field
// elements per leaf in struct order, each the canonical 32-byte little-endian `to_repr()` bytes whatever the context's own
// format — behind a fixed 64-byte header, followed by the root. The levels are not stored: loading re-hashes the leaves on the
// GPU (0.53 s at depth 24) and REFUSES the file if the rebuilt root differs from the stored one.
//
//   offset  0  char[8]  "IMTB200\0"
//           8  u32      version (1)
//          12  u32      t, 16 rate, 20 r_f, 24 r_p        the Poseidon instance the root was computed with
//          28  u32      depth
//          32  u64      n = number of leaves
//          40  u8[24]   zero
//          64  n x 96   leaves: val, next_val, next_idx  (32-byte LE canonical each)
//    64 + 96n  u8[32]   root (32-byte LE canonical)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "imt_b200.h"
#include "imt_internal.h"

using namespace imt;
using namespace imt_host;

namespace {

constexpr char kMagic[8] = {'I', 'M', 'T', 'B', '2', '0', '0', '\0'};
constexpr size_t kChunkLeaves = (size_t)1 << 19;  // 48 MiB per transfer

struct File {
    std::FILE* f = nullptr;
    ~File() {
        if (f) std::fclose(f);
    }
};

void put_header(uint8_t* h, const imt_ctx* ctx, unsigned depth, uint64_t n) {
    std::memset(h, 0, 64);
    std::memcpy(h, kMagic, 8);
    const uint32_t w[6] = {1u, ctx->spec.t, ctx->spec.t - 1, ctx->spec.r_f, ctx->spec.r_p, depth};
    std::memcpy(h + 8, w, sizeof(w));
    std::memcpy(h + 32, &n, 8);
}

imt_status parse_header(const uint8_t* h, imt_checkpoint_info* info) {
    if (std::memcmp(h, kMagic, 8) != 0) return IMT_ERR_INVALID_ARG;
    uint32_t w[6];
    std::memcpy(w, h + 8, sizeof(w));
    uint64_t n;
    std::memcpy(&n, h + 32, 8);
    if (w[0] != 1 || n == 0) return IMT_ERR_INVALID_ARG;
    info->version = w[0], info->t = w[1], info->rate = w[2], info->r_f = w[3], info->r_p = w[4], info->depth = w[5], info->num_leaves = n;
    return IMT_OK;
}

// shards of one tree in rank order -> one flat file
imt_status save_shards(const std::vector<imt_tree*>& shards, const char* path) {
    imt_ctx* c0 = shards[0]->ctx;
    if (!path) return fail(c0, IMT_ERR_INVALID_ARG, "null path");
    for (imt_tree* t : shards)
        if (!t->d_pre) return fail(c0, IMT_ERR_INVALID_ARG, "tree was not built from leaves: nothing to checkpoint");
    const uint64_t n_total = (uint64_t)shards[0]->n * shards.size();
    File out;
    out.f = std::fopen(path, "wb");
    if (!out.f) return fail(c0, IMT_ERR_INVALID_ARG, "cannot open the checkpoint file for writing");
    uint8_t header[64];
    put_header(header, c0, imt_tree_depth(shards[0]), n_total);
    if (std::fwrite(header, 1, 64, out.f) != 64) return fail(c0, IMT_ERR_INVALID_ARG, "short write");
    std::vector<Fr> host(3 * std::min<size_t>(kChunkLeaves, shards[0]->n));
    for (imt_tree* t : shards) {
        imt_ctx* ctx = t->ctx;
        IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
        DevBuf tmp(ctx);
        if (ctx->fmt != kFmtCanonical) IMT_TRY_CUDA(ctx, tmp.alloc(host.size() * sizeof(Fr)));
        for (size_t off = 0; off < t->n; off += kChunkLeaves) {
            const size_t cnt = std::min(kChunkLeaves, t->n - off);
            const Fr* src = t->d_pre + 3 * off;
            if (ctx->fmt != kFmtCanonical) {
                IMT_TRY(clear_err(ctx));
                IMT_TRY(launch_convert(ctx, src, tmp.p, 3 * cnt, ctx->fmt, kFmtCanonical));
                src = tmp.as<Fr>();
            }
            IMT_TRY_CUDA(ctx, cudaMemcpyAsync(host.data(), src, 3 * cnt * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
            IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (std::fwrite(host.data(), sizeof(Fr), 3 * cnt, out.f) != 3 * cnt) return fail(c0, IMT_ERR_INVALID_ARG, "short write");
        }
    }
    // the root in canonical bytes: through a canonical-format read of the root element
    Fr root;
    {
        imt_tree* t = shards[0];
        imt_ctx* ctx = t->ctx;
        IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
        const Fr* d_root = t->cap_valid ? t->d_cap + level_offset(t->world, t->cap_depth) : t->d_levels + level_offset(t->n, t->depth);
        DevBuf tmp(ctx);
        IMT_TRY_CUDA(ctx, tmp.alloc(sizeof(Fr)));
        IMT_TRY(clear_err(ctx));
        IMT_TRY(launch_convert(ctx, d_root, tmp.p, 1, kFmtMontgomery, kFmtCanonical));
        IMT_TRY_CUDA(ctx, cudaMemcpyAsync(&root, tmp.p, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
        IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (std::fwrite(&root, sizeof(Fr), 1, out.f) != 1 || std::fflush(out.f) != 0) return fail(c0, IMT_ERR_INVALID_ARG, "short write");
    return IMT_OK;
}

// reads the leaves of [first, first + n) of an open checkpoint into a tree's preimage buffer (converted to the context format)
imt_status load_leaves(std::FILE* f, imt_tree* t, uint64_t first) {
    imt_ctx* ctx = t->ctx;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<Fr> host(3 * std::min<size_t>(kChunkLeaves, t->n));
    DevBuf tmp(ctx);
    IMT_TRY_CUDA(ctx, tmp.alloc(host.size() * sizeof(Fr)));
    if (std::fseek(f, (long)(64 + first * 96), SEEK_SET) != 0) return fail(ctx, IMT_ERR_INVALID_ARG, "checkpoint is truncated");
    IMT_TRY(clear_err(ctx));
    for (size_t off = 0; off < t->n; off += kChunkLeaves) {
        const size_t cnt = std::min(kChunkLeaves, t->n - off);
        if (std::fread(host.data(), sizeof(Fr), 3 * cnt, f) != 3 * cnt) return fail(ctx, IMT_ERR_INVALID_ARG, "checkpoint is truncated");
        IMT_TRY_CUDA(ctx, cudaMemcpyAsync(tmp.p, host.data(), 3 * cnt * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
        IMT_TRY(launch_convert(ctx, tmp.p, t->d_pre + 3 * off, 3 * cnt, kFmtCanonical, ctx->fmt));  // validates every element < p
        IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));                                     // host[] is reused
    }
    return finish(ctx);
}

imt_status open_checkpoint(imt_ctx* ctx, const char* path, File* in, imt_checkpoint_info* info) {
    if (!path) return fail(ctx, IMT_ERR_INVALID_ARG, "null path");
    in->f = std::fopen(path, "rb");
    if (!in->f) return fail(ctx, IMT_ERR_INVALID_ARG, "cannot open the checkpoint file");
    uint8_t header[64];
    if (std::fread(header, 1, 64, in->f) != 64 || parse_header(header, info) != IMT_OK) return fail(ctx, IMT_ERR_INVALID_ARG, "not an imt_b200 checkpoint");
    if (info->t != ctx->spec.t || info->r_f != ctx->spec.r_f || info->r_p != ctx->spec.r_p)
        return fail(ctx, IMT_ERR_INVALID_ARG, "the checkpoint was written with another Poseidon instance");
    if (std::fseek(in->f, (long)(64 + info->num_leaves * 96), SEEK_SET) != 0 || std::fread(info->root, 1, 32, in->f) != 32)
        return fail(ctx, IMT_ERR_INVALID_ARG, "checkpoint is truncated");
    return IMT_OK;
}

imt_status root_matches(imt_tree* t, const uint8_t* stored) {
    imt_ctx* ctx = t->ctx;
    const Fr* d_root = t->cap_valid ? t->d_cap + level_offset(t->world, t->cap_depth) : t->d_levels + level_offset(t->n, t->depth);
    DevBuf tmp(ctx);
    Fr root;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    IMT_TRY_CUDA(ctx, tmp.alloc(sizeof(Fr)));
    IMT_TRY(clear_err(ctx));
    IMT_TRY(launch_convert(ctx, d_root, tmp.p, 1, kFmtMontgomery, kFmtCanonical));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(&root, tmp.p, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (std::memcmp(&root, stored, 32) != 0) return fail(ctx, IMT_ERR_INVALID_ARG, "checkpoint is corrupt: the rebuilt root differs from the stored one");
    return IMT_OK;
}

}  // namespace

extern "C" imt_status imt_checkpoint_read_info(const char* path, imt_checkpoint_info* info) {
    if (!path || !info) return IMT_ERR_INVALID_ARG;
    File in;
    in.f = std::fopen(path, "rb");
    if (!in.f) return IMT_ERR_INVALID_ARG;
    uint8_t header[64];
    if (std::fread(header, 1, 64, in.f) != 64) return IMT_ERR_INVALID_ARG;
    IMT_TRY(parse_header(header, info));
    if (std::fseek(in.f, (long)(64 + info->num_leaves * 96), SEEK_SET) != 0 || std::fread(info->root, 1, 32, in.f) != 32) return IMT_ERR_INVALID_ARG;
    return IMT_OK;
}

extern "C" imt_status imt_tree_save(imt_tree* tree, const char* path) {
    if (!tree) return IMT_ERR_INVALID_ARG;
    if (tree->world > 1) return fail(tree->ctx, IMT_ERR_INVALID_ARG, "a shard cannot be checkpointed alone: use imt_mtree_save, or gather the leaves");
    return save_shards({tree}, path);
}

extern "C" imt_status imt_tree_load(imt_ctx* ctx, const char* path, imt_tree** out) {
    if (!ctx || !out) return IMT_ERR_INVALID_ARG;
    *out = nullptr;
    File in;
    imt_checkpoint_info info;
    IMT_TRY(open_checkpoint(ctx, path, &in, &info));
    IMT_TRY(check_leaf_count(ctx, info.num_leaves));
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    imt_tree* t = nullptr;
    IMT_TRY(tree_alloc(ctx, info.num_leaves, true, &t));
    imt_status st = load_leaves(in.f, t, 0);
    if (st == IMT_OK) st = enqueue_rebuild(t, t->d_pre, true);
    if (st == IMT_OK) st = finish(ctx);
    if (st == IMT_OK) st = root_matches(t, info.root);
    if (st != IMT_OK) {
        cudaStreamSynchronize(ctx->stream);
        imt_tree_destroy(t);
        return st;
    }
    *out = t;
    return IMT_OK;
}

extern "C" imt_status imt_mtree_save(imt_mtree* mt, const char* path) {
    imt_group* g = nullptr;
    std::vector<imt_tree*> shards;
    IMT_TRY(mtree_parts(mt, &g, &shards));
    return save_shards(shards, path);
}

extern "C" imt_status imt_multi_load(imt_multi* m, const char* path, imt_mtree** out) {
    if (!m || !out) return IMT_ERR_INVALID_ARG;
    *out = nullptr;
    imt_ctx* c0 = imt_multi_ctx(m, 0);
    const unsigned world = imt_multi_size(m);
    File in;
    imt_checkpoint_info info;
    IMT_TRY(open_checkpoint(c0, path, &in, &info));
    IMT_TRY(check_leaf_count(c0, info.num_leaves));
    if (info.num_leaves < world) return fail(c0, IMT_ERR_INVALID_ARG, "fewer leaves than devices");
    // the leaves are streamed from the file shard by shard, then every device re-hashes its resident preimages
    imt_mtree* mt = nullptr;
    IMT_TRY(mtree_alloc(m, info.num_leaves, &mt));
    imt_status st = IMT_OK;
    for (unsigned i = 0; i < world && st == IMT_OK; ++i) st = load_leaves(in.f, imt_mtree_shard(mt, i), (uint64_t)i * (info.num_leaves / world));
    if (st == IMT_OK) st = mtree_rebuild_resident(mt);
    if (st == IMT_OK) st = root_matches(imt_mtree_shard(mt, 0), info.root);
    if (st != IMT_OK) {
        imt_mtree_destroy(mt);
        return st;
    }
    *out = mt;
    return IMT_OK;
}
