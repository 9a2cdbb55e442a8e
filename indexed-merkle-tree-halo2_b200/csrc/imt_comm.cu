// Multi-GPU behind the C-ABI: NCCL lives INSIDE libimt_b200.so, so a Rust / C host reaches the sharded build with one call
// and no Python. Replaces the root exchange the reference does not have (its tree is one Vec<Vec<F>> on one CPU thread,
// /root/reference/src/utils.rs:20-57): a depth-d tree over N = 2^k GPUs is N independent depth-(d-k) subtrees plus a k-level
// cap, so `IndexedMerkleTree::new` becomes  N local builds -> ONE ncclAllGather of N x 32 bytes over NVLink -> k cap levels.
//
// Two ways to form the group of ranks:
//   imt_comm_create    one process per GPU (torchrun / MPI layout): rank 0 makes an id (imt_comm_unique_id), every rank
//                      attaches a communicator to its own context with ncclCommInitRank
//   imt_multi_create   one process drives all N GPUs: N contexts + ncclCommInitAll; every collective is a
//                      ncclGroupStart / ncclGroupEnd bracket over the N communicators
// NCCL is bound at run time (dlopen of libnccl.so.2 on the first communicator, reusing the copy the process has already
// loaded — e.g. torch's): single-GPU users need no NCCL at all, and loading this library never changes which NCCL a later
// `import torch` resolves. If a device is listed twice in imt_multi_create (tests on a one-GPU box) NCCL cannot be used
// (it refuses duplicate devices) and the group falls back to device-to-device copies on the streams — the rest of the
// code path is identical.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "imt_b200.h"
#include "imt_internal.h"

using namespace imt;
using namespace imt_host;

// ------------------------------------------------------------------------------------------------- NCCL, bound at run time
namespace {

struct NcclApi {
    void* handle = nullptr;
    std::string error;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi* nccl(std::string* why) {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = std::getenv("IMT_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy this process already has (torch's)
            if (!api.handle) api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("libnccl.so.2 could not be loaded: ") + (dlerror() ? dlerror() : "not found");
            return;
        }
        auto bind = [&](auto& fn, const char* sym) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.handle, sym));
            if (!fn && api.error.empty()) api.error = std::string("libnccl has no symbol ") + sym;
        };
        bind(api.GetVersion, "ncclGetVersion");
        bind(api.GetUniqueId, "ncclGetUniqueId");
        bind(api.CommInitRank, "ncclCommInitRank");
        bind(api.CommInitAll, "ncclCommInitAll");
        bind(api.CommDestroy, "ncclCommDestroy");
        bind(api.AllGather, "ncclAllGather");
        bind(api.AllReduce, "ncclAllReduce");
        bind(api.GroupStart, "ncclGroupStart");
        bind(api.GroupEnd, "ncclGroupEnd");
        bind(api.GetErrorString, "ncclGetErrorString");
    });
    if (!api.error.empty()) {
        if (why) *why = api.error;
        return nullptr;
    }
    return &api;
}

}  // namespace

// ------------------------------------------------------------------------------------------------- groups
struct imt_group {
    unsigned world = 1;
    bool use_nccl = true;         // false: every rank is local and at least one device appears twice -> stream-ordered copies
    bool owns_ctxs = false;       // imt_multi_create made the contexts
    std::vector<imt_ctx*> ctxs;   // local contexts, ascending rank
    std::vector<unsigned> ranks;  // their ranks
    std::vector<ncclComm_t> comms;
    std::vector<cudaEvent_t> ready, done;  // copy transport: per local rank "send buffer written" / "my copies finished"
    int nccl_version = 0;
    std::string last_error;
};
struct imt_multi {
    imt_group g;
};
struct imt_mtree {
    imt_multi* m = nullptr;
    std::vector<imt_tree*> shards;  // shard i on device i, rank i
    size_t n_total = 0;
};

namespace {

imt_status group_fail(imt_group* g, const char* what, const char* detail) {
    g->last_error = std::string(what) + ": " + (detail ? detail : "");
    for (imt_ctx* c : g->ctxs) c->last_error = g->last_error;
    return IMT_ERR_CUDA;
}
#define IMT_TRY_NCCL(g, api, expr)                                                         \
    do {                                                                                   \
        ncclResult_t r_ = (expr);                                                          \
        if (r_ != ncclSuccess) return group_fail(g, #expr, (api)->GetErrorString(r_));     \
    } while (0)
#define IMT_TRY_CUDA_G(g, expr)                                                            \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess) return group_fail(g, #expr, cudaGetErrorString(e_));        \
    } while (0)

void group_release(imt_group* g) {
    std::string why;
    const NcclApi* api = g->comms.empty() ? nullptr : nccl(&why);
    for (size_t i = 0; i < g->ctxs.size(); ++i) {
        cudaSetDevice(g->ctxs[i]->device);
        cudaStreamSynchronize(g->ctxs[i]->stream);
        if (api && i < g->comms.size() && g->comms[i]) api->CommDestroy(g->comms[i]);
        if (i < g->ready.size() && g->ready[i]) cudaEventDestroy(g->ready[i]);
        if (i < g->done.size() && g->done[i]) cudaEventDestroy(g->done[i]);
        g->ctxs[i]->group = nullptr;
    }
    g->comms.clear();
    g->ready.clear();
    g->done.clear();
}

// sum over the `world` rank-major slices of `all` (copy transport only)
__global__ void k_sum_slices_u64(const unsigned long long* __restrict__ all, unsigned world, size_t count, unsigned long long* __restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long s = 0;
    for (unsigned r = 0; r < world; ++r) s += all[(size_t)r * count + i];
    out[i] = s;
}
__global__ void k_sum_slices_u8(const uint8_t* __restrict__ all, unsigned world, size_t count, uint8_t* __restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned s = 0;
    for (unsigned r = 0; r < world; ++r) s += all[(size_t)r * count + i];
    out[i] = (uint8_t)s;
}

// copy transport: all ranks are local. Stream i waits until every send buffer is written, pulls the N pieces, and every
// producer waits for all pulls before it may touch its send buffer again.
imt_status copy_all_gather(imt_group* g, const std::vector<const void*>& send, const std::vector<void*>& recv, size_t bytes) {
    const size_t n = g->ctxs.size();
    for (size_t j = 0; j < n; ++j) {
        IMT_TRY_CUDA_G(g, cudaSetDevice(g->ctxs[j]->device));
        IMT_TRY_CUDA_G(g, cudaEventRecord(g->ready[j], g->ctxs[j]->stream));
    }
    for (size_t i = 0; i < n; ++i) {
        IMT_TRY_CUDA_G(g, cudaSetDevice(g->ctxs[i]->device));
        for (size_t j = 0; j < n; ++j) {
            if (j != i) IMT_TRY_CUDA_G(g, cudaStreamWaitEvent(g->ctxs[i]->stream, g->ready[j], 0));
            char* dst = (char*)recv[i] + (size_t)g->ranks[j] * bytes;
            if (g->ctxs[i]->device == g->ctxs[j]->device)
                IMT_TRY_CUDA_G(g, cudaMemcpyAsync(dst, send[j], bytes, cudaMemcpyDeviceToDevice, g->ctxs[i]->stream));
            else  // explicit peer copy: pool memory of another device is not addressable through the unified-addressing default kind
                IMT_TRY_CUDA_G(g, cudaMemcpyPeerAsync(dst, g->ctxs[i]->device, send[j], g->ctxs[j]->device, bytes, g->ctxs[i]->stream));
        }
        IMT_TRY_CUDA_G(g, cudaEventRecord(g->done[i], g->ctxs[i]->stream));
    }
    for (size_t j = 0; j < n; ++j) {
        IMT_TRY_CUDA_G(g, cudaSetDevice(g->ctxs[j]->device));
        for (size_t i = 0; i < n; ++i)
            if (i != j) IMT_TRY_CUDA_G(g, cudaStreamWaitEvent(g->ctxs[j]->stream, g->done[i], 0));
    }
    return IMT_OK;
}

}  // namespace

namespace imt_host {

unsigned group_world(const imt_group* g) { return g->world; }
unsigned group_local_count(const imt_group* g) { return (unsigned)g->ctxs.size(); }
imt_ctx* group_ctx(imt_group* g, unsigned slot) { return g->ctxs[slot]; }
unsigned group_rank(const imt_group* g, unsigned slot) { return g->ranks[slot]; }

imt_status group_all_gather(imt_group* g, const std::vector<const void*>& send, const std::vector<void*>& recv, size_t bytes) {
    if (bytes == 0) return IMT_OK;
    if (!g->use_nccl) return copy_all_gather(g, send, recv, bytes);
    std::string why;
    const NcclApi* api = nccl(&why);
    if (!api) return group_fail(g, "NCCL", why.c_str());
    IMT_TRY_NCCL(g, api, api->GroupStart());
    for (size_t i = 0; i < g->ctxs.size(); ++i) {
        ncclResult_t r = api->AllGather(send[i], recv[i], bytes, ncclChar, g->comms[i], g->ctxs[i]->stream);
        if (r != ncclSuccess) {
            api->GroupEnd();
            return group_fail(g, "ncclAllGather", api->GetErrorString(r));
        }
        ++g->ctxs[i]->launches;
    }
    IMT_TRY_NCCL(g, api, api->GroupEnd());
    return IMT_OK;
}

imt_status group_all_reduce_sum(imt_group* g, const std::vector<void*>& buf, size_t count, bool bytes8) {
    if (count == 0 || g->world == 1) return IMT_OK;
    if (g->use_nccl) {
        std::string why;
        const NcclApi* api = nccl(&why);
        if (!api) return group_fail(g, "NCCL", why.c_str());
        IMT_TRY_NCCL(g, api, api->GroupStart());
        for (size_t i = 0; i < g->ctxs.size(); ++i) {
            ncclResult_t r = api->AllReduce(buf[i], buf[i], count, bytes8 ? ncclUint8 : ncclUint64, ncclSum, g->comms[i], g->ctxs[i]->stream);
            if (r != ncclSuccess) {
                api->GroupEnd();
                return group_fail(g, "ncclAllReduce", api->GetErrorString(r));
            }
            ++g->ctxs[i]->launches;
        }
        IMT_TRY_NCCL(g, api, api->GroupEnd());
        return IMT_OK;
    }
    // copy transport: gather every rank's buffer, then sum the slices locally
    const size_t n = g->ctxs.size(), bytes = count * (bytes8 ? 1 : 8);
    std::vector<DevBuf*> tmp(n, nullptr);
    std::vector<const void*> send(n);
    std::vector<void*> recv(n);
    imt_status st = IMT_OK;
    for (size_t i = 0; i < n && st == IMT_OK; ++i) {
        cudaSetDevice(g->ctxs[i]->device);
        tmp[i] = new DevBuf(g->ctxs[i]);
        if (tmp[i]->alloc(bytes * g->world) != cudaSuccess) st = group_fail(g, "cudaMallocAsync", "all-reduce scratch");
        send[i] = buf[i];
        recv[i] = tmp[i]->p;
    }
    if (st == IMT_OK) st = copy_all_gather(g, send, recv, bytes);
    for (size_t i = 0; i < n && st == IMT_OK; ++i) {
        cudaSetDevice(g->ctxs[i]->device);
        if (bytes8) k_sum_slices_u8<<<grid_for(count, 256), 256, 0, g->ctxs[i]->stream>>>((const uint8_t*)recv[i], g->world, count, (uint8_t*)buf[i]);
        else k_sum_slices_u64<<<grid_for(count, 256), 256, 0, g->ctxs[i]->stream>>>((const unsigned long long*)recv[i], g->world, count, (unsigned long long*)buf[i]);
        ++g->ctxs[i]->launches;
    }
    for (size_t i = 0; i < n; ++i) {
        if (tmp[i]) cudaSetDevice(g->ctxs[i]->device);
        delete tmp[i];  // stream-ordered free
    }
    return st;
}

}  // namespace imt_host

// ------------------------------------------------------------------------------------------------- one process per GPU
extern "C" imt_status imt_comm_unique_id(void* id) {
    if (!id) return IMT_ERR_INVALID_ARG;
    std::string why;
    const NcclApi* api = nccl(&why);
    if (!api) {
        std::fprintf(stderr, "imt_comm_unique_id: %s\n", why.c_str());
        return IMT_ERR_CUDA;
    }
    ncclUniqueId uid;
    if (api->GetUniqueId(&uid) != ncclSuccess) return IMT_ERR_CUDA;
    static_assert(sizeof(uid) == IMT_COMM_ID_BYTES, "IMT_COMM_ID_BYTES must equal NCCL_UNIQUE_ID_BYTES");
    std::memcpy(id, &uid, sizeof(uid));
    return IMT_OK;
}

extern "C" imt_status imt_comm_create(imt_ctx* ctx, unsigned rank, unsigned world, const void* id) {
    if (!ctx) return IMT_ERR_INVALID_ARG;
    if (!id || world == 0 || (world & (world - 1)) || rank >= world) return fail(ctx, IMT_ERR_INVALID_ARG, "bad rank/world (world must be a power of two)");
    if (ctx->group) return fail(ctx, IMT_ERR_INVALID_ARG, "the context already belongs to a group");
    std::string why;
    const NcclApi* api = nccl(&why);
    if (!api) return fail(ctx, IMT_ERR_CUDA, why.c_str());
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    imt_group* g = new (std::nothrow) imt_group();
    if (!g) return fail(ctx, IMT_ERR_CUDA, "out of host memory");
    g->world = world;
    g->ctxs = {ctx};
    g->ranks = {rank};
    g->comms = {nullptr};
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    ncclResult_t r = api->CommInitRank(&g->comms[0], (int)world, uid, (int)rank);
    if (r != ncclSuccess) {
        ctx->last_error = std::string("ncclCommInitRank: ") + api->GetErrorString(r);
        delete g;
        return IMT_ERR_CUDA;
    }
    api->GetVersion(&g->nccl_version);
    ctx->group = g;
    ctx->group_slot = 0;
    return IMT_OK;
}

extern "C" imt_status imt_comm_destroy(imt_ctx* ctx) {
    if (!ctx) return IMT_ERR_INVALID_ARG;
    imt_group* g = ctx->group;
    if (!g) return IMT_OK;
    if (g->owns_ctxs) return fail(ctx, IMT_ERR_INVALID_ARG, "this context belongs to an imt_multi: destroy that instead");
    group_release(g);
    delete g;
    return IMT_OK;
}

extern "C" imt_status imt_comm_info(const imt_ctx* ctx, unsigned* rank, unsigned* world, int* nccl_version) {
    if (!ctx) return IMT_ERR_INVALID_ARG;
    const imt_group* g = ctx->group;
    if (rank) *rank = g ? g->ranks[ctx->group_slot] : 0;
    if (world) *world = g ? g->world : 1;
    if (nccl_version) *nccl_version = g ? g->nccl_version : 0;
    return IMT_OK;
}

// ------------------------------------------------------------------------------------------------- root exchange
namespace {

// The N subtree roots -> level 0 of every rank's cap, then the cap levels. Nothing is converted or staged: a subtree root
// already sits in Montgomery form at the end of d_levels (the last kernel of the local build wrote it there), and that slot
// IS the all-gather's send buffer; the receive buffer IS level 0 of the cap.
imt_status exchange_roots(imt_group* g, const std::vector<imt_tree*>& trees) {
    const unsigned world = g->world;
    unsigned cap_depth = 0;
    while ((1u << cap_depth) < world) ++cap_depth;
    std::vector<const void*> send(trees.size());
    std::vector<void*> recv(trees.size());
    for (size_t i = 0; i < trees.size(); ++i) {
        imt_tree* t = trees[i];
        imt_ctx* ctx = t->ctx;
        if (t->n != trees[0]->n) return fail(ctx, IMT_ERR_INVALID_ARG, "every rank must hold the same number of leaves");
        IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
        t->cap_valid = false;
        if (t->cap_alloc_world != world) {
            if (t->d_cap) tree_free(ctx, t->d_cap), t->d_cap = nullptr;
            IMT_TRY_CUDA(ctx, tree_malloc(ctx, (void**)&t->d_cap, (2 * (size_t)world - 1) * sizeof(Fr)));
            t->cap_alloc_world = world;
        }
        const unsigned rank = g->ranks[ctx->group_slot];
        if (t->rank != rank || t->world != world) invalidate_index(t);  // slot numbers of the index are global
        t->rank = rank;
        t->world = world;
        t->cap_depth = cap_depth;
        send[i] = t->d_levels + level_offset(t->n, t->depth);
        recv[i] = t->d_cap;
    }
    IMT_TRY(group_all_gather(g, send, recv, sizeof(Fr)));
    for (imt_tree* t : trees) {
        imt_ctx* ctx = t->ctx;
        IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
        for (unsigned l = 0; l < cap_depth; ++l)
            IMT_TRY(launch_level(ctx, t->d_cap + level_offset(world, l), t->d_cap + level_offset(world, l + 1), world >> (l + 1)));
    }
    for (imt_tree* t : trees) {
        IMT_TRY(finish(t->ctx));
        t->cap_valid = true;
    }
    return IMT_OK;
}

imt_status tree_group(imt_tree* t, imt_group** g) {
    if (!t) return IMT_ERR_INVALID_ARG;
    if (!t->ctx->group) return fail(t->ctx, IMT_ERR_INVALID_ARG, "the context has no communicator: call imt_comm_create (or use imt_multi_create) first");
    if (t->ctx->group->owns_ctxs) return fail(t->ctx, IMT_ERR_INVALID_ARG, "this tree is a shard of an imt_mtree: use the imt_mtree_* calls");
    *g = t->ctx->group;
    return IMT_OK;
}

}  // namespace

extern "C" imt_status imt_tree_exchange_roots(imt_tree* tree) {
    imt_group* g = nullptr;
    IMT_TRY(tree_group(tree, &g));
    return exchange_roots(g, {tree});
}

static imt_status sharded_build(imt_ctx* ctx, const void* preimages, size_t n_local, bool device_src, imt_tree** out) {
    if (!ctx || !out) return IMT_ERR_INVALID_ARG;
    *out = nullptr;
    if (!ctx->group || ctx->group->owns_ctxs) return fail(ctx, IMT_ERR_INVALID_ARG, "the context has no communicator: call imt_comm_create first");
    IMT_TRY(check_leaf_count(ctx, n_local));
    if (!preimages) return fail(ctx, IMT_ERR_INVALID_ARG, "null preimages");
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    imt_tree* t = nullptr;
    IMT_TRY(tree_alloc(ctx, n_local, true, &t));
    imt_status st = enqueue_rebuild(t, preimages, device_src);  // the all-gather below is stream-ordered behind the last level
    if (st == IMT_OK && !device_src) st = wait_staging(ctx);
    if (st == IMT_OK) st = exchange_roots(ctx->group, {t});
    if (st != IMT_OK) {
        cudaStreamSynchronize(ctx->stream);
        imt_tree_destroy(t);
        return st;
    }
    *out = t;
    return IMT_OK;
}
extern "C" imt_status imt_sharded_build_from_leaves(imt_ctx* ctx, const void* local_preimages, size_t n_local, imt_tree** out) {
    return sharded_build(ctx, local_preimages, n_local, false, out);
}
extern "C" imt_status imt_sharded_build_from_leaves_dev(imt_ctx* ctx, const void* d_local_preimages, size_t n_local, imt_tree** out) {
    return sharded_build(ctx, d_local_preimages, n_local, true, out);
}
static imt_status sharded_rebuild(imt_tree* t, const void* preimages, bool device_src) {
    imt_group* g = nullptr;
    IMT_TRY(tree_group(t, &g));
    imt_status st = enqueue_rebuild(t, preimages, device_src);
    if (st == IMT_OK && !device_src) st = wait_staging(t->ctx);
    if (st != IMT_OK) {
        cudaStreamSynchronize(t->ctx->stream);
        return st;
    }
    return exchange_roots(g, {t});
}
extern "C" imt_status imt_sharded_rebuild_from_leaves(imt_tree* t, const void* local_preimages) { return sharded_rebuild(t, local_preimages, false); }
extern "C" imt_status imt_sharded_rebuild_from_leaves_dev(imt_tree* t, const void* d_local_preimages) { return sharded_rebuild(t, d_local_preimages, true); }

// ------------------------------------------------------------------------------------------------- one process, N GPUs
extern "C" imt_status imt_multi_create(const int* devices, unsigned n_dev, imt_fe_format format, imt_multi** out) {
    if (!out) return IMT_ERR_INVALID_ARG;
    *out = nullptr;
    if (!devices || n_dev == 0 || (n_dev & (n_dev - 1)) || n_dev > 64) return IMT_ERR_INVALID_ARG;
    imt_multi* m = new (std::nothrow) imt_multi();
    if (!m) return IMT_ERR_CUDA;
    imt_group* g = &m->g;
    g->world = n_dev;
    g->owns_ctxs = true;
    bool duplicate = false;
    for (unsigned i = 0; i < n_dev; ++i)
        for (unsigned j = 0; j < i; ++j) duplicate |= devices[i] == devices[j];
    g->use_nccl = !duplicate && n_dev > 1 && !std::getenv("IMT_MULTI_NO_NCCL");
    imt_status st = IMT_OK;
    for (unsigned i = 0; i < n_dev && st == IMT_OK; ++i) {
        imt_ctx* c = nullptr;
        st = imt_ctx_create(devices[i], format, &c);
        if (st == IMT_OK) {
            c->group = g;
            c->group_slot = i;
            g->ctxs.push_back(c);
            g->ranks.push_back(i);
        }
    }
    if (st == IMT_OK && g->use_nccl) {
        std::string why;
        const NcclApi* api = nccl(&why);
        if (!api) {
            std::fprintf(stderr, "imt_multi_create: %s\n", why.c_str());
            st = IMT_ERR_CUDA;
        } else {
            g->comms.assign(n_dev, nullptr);
            ncclResult_t r = api->CommInitAll(g->comms.data(), (int)n_dev, devices);
            if (r != ncclSuccess) {
                std::fprintf(stderr, "imt_multi_create: ncclCommInitAll: %s\n", api->GetErrorString(r));
                g->comms.clear();
                st = IMT_ERR_CUDA;
            } else {
                api->GetVersion(&g->nccl_version);
            }
        }
    }
    if (st == IMT_OK && !g->use_nccl) {
        g->ready.assign(n_dev, nullptr);
        g->done.assign(n_dev, nullptr);
        for (unsigned i = 0; i < n_dev && st == IMT_OK; ++i) {
            cudaSetDevice(devices[i]);
            if (cudaEventCreateWithFlags(&g->ready[i], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&g->done[i], cudaEventDisableTiming) != cudaSuccess)
                st = IMT_ERR_CUDA;
        }
    }
    if (st == IMT_OK) {  // peer access: kernels of one device may read the stored levels of another (query-sharded paths / traces)
        for (unsigned i = 0; i < n_dev; ++i)
            for (unsigned j = 0; j < n_dev; ++j) {
                if (devices[i] == devices[j]) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
                if (can) {
                    cudaSetDevice(devices[i]);
                    cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                    if (e != cudaSuccess) cudaGetLastError();  // already enabled
                    // ... and the buffers of context j (its private pool) become addressable from device i
                    cudaMemAccessDesc desc = {};
                    desc.location.type = cudaMemLocationTypeDevice;
                    desc.location.id = devices[i];
                    desc.flags = cudaMemAccessFlagsProtReadWrite;
                    if (cudaMemPoolSetAccess(g->ctxs[j]->pool, &desc, 1) != cudaSuccess) cudaGetLastError();
                }
            }
    }
    if (st != IMT_OK) {
        imt_multi_destroy(m);
        return st;
    }
    *out = m;
    return IMT_OK;
}

extern "C" void imt_multi_destroy(imt_multi* m) {
    if (!m) return;
    std::vector<imt_ctx*> ctxs = m->g.ctxs;
    group_release(&m->g);
    for (imt_ctx* c : ctxs) imt_ctx_destroy(c);
    delete m;
}
extern "C" unsigned imt_multi_size(const imt_multi* m) { return m ? m->g.world : 0; }
extern "C" imt_ctx* imt_multi_ctx(imt_multi* m, unsigned i) { return (m && i < m->g.ctxs.size()) ? m->g.ctxs[i] : nullptr; }
extern "C" const char* imt_multi_last_error(const imt_multi* m) {
    if (!m) return "null imt_multi";
    for (const imt_ctx* c : m->g.ctxs)
        if (!c->last_error.empty()) return c->last_error.c_str();
    return m->g.last_error.c_str();
}
extern "C" int imt_multi_uses_nccl(const imt_multi* m) { return m && m->g.use_nccl ? m->g.nccl_version : 0; }

namespace {
// preimages == nullptr: re-hash the leaves already resident in every shard's preimage buffer (checkpoint load)
imt_status mtree_rebuild(imt_mtree* mt, const void* preimages) {
    imt_group* g = &mt->m->g;
    const size_t n_local = mt->n_total / g->world;
    // queue every device's pipeline first (H2D chunks + leaf kernels + levels are all asynchronous), THEN wait: the N builds overlap.
    // Host leaves: one host thread per device. From page-locked memory the copies are asynchronous and one thread would do, but from
    // pageable memory (a Rust Vec<F>) every cudaMemcpyAsync returns only after its chunk is staged — one thread would feed the devices
    // one after the other (device 7 starting 7 x 190 MB of staging late at depth 24) instead of all links at once.
    imt_status st = IMT_OK;
    if (preimages && g->world > 1) {
        std::vector<imt_status> sts(g->world, IMT_OK);
        std::vector<std::thread> th;
        th.reserve(g->world);
        for (unsigned i = 0; i < g->world; ++i)
            th.emplace_back([&, i] {
                sts[i] = enqueue_rebuild(mt->shards[i], static_cast<const char*>(preimages) + (size_t)i * n_local * 3 * sizeof(Fr), false);
                const imt_status s2 = wait_staging(g->ctxs[i]);  // also after a failed enqueue: nothing may still read the caller's buffer
                if (sts[i] == IMT_OK) sts[i] = s2;
            });
        for (auto& x : th) x.join();
        for (unsigned i = 0; i < g->world; ++i)
            if (sts[i] != IMT_OK && st == IMT_OK) {
                st = sts[i];
                if (i) g->ctxs[0]->last_error = g->ctxs[i]->last_error;
            }
    } else {
        for (unsigned i = 0; i < g->world && st == IMT_OK; ++i) {
            if (preimages) st = enqueue_rebuild(mt->shards[i], static_cast<const char*>(preimages) + (size_t)i * n_local * 3 * sizeof(Fr), false);
            else st = enqueue_rebuild(mt->shards[i], mt->shards[i]->d_pre, true);
        }
        for (unsigned i = 0; i < g->world && preimages; ++i) {
            const imt_status s2 = wait_staging(g->ctxs[i]);
            if (st == IMT_OK) st = s2;
        }
    }
    if (st == IMT_OK) st = exchange_roots(g, mt->shards);
    if (st != IMT_OK)
        for (imt_ctx* c : g->ctxs) {
            cudaSetDevice(c->device);
            cudaStreamSynchronize(c->stream);
        }
    return st;
}
}  // namespace

namespace imt_host {
// an imt_mtree of n leaves with every buffer allocated and nothing built (imt_multi_build_from_leaves, checkpoint load)
imt_status mtree_alloc(imt_multi* m, size_t n, imt_mtree** out) {
    imt_group* g = &m->g;
    imt_ctx* c0 = g->ctxs[0];
    imt_mtree* mt = new (std::nothrow) imt_mtree();
    if (!mt) return fail(c0, IMT_ERR_CUDA, "out of host memory");
    mt->m = m;
    mt->n_total = n;
    for (unsigned i = 0; i < g->world; ++i) {
        cudaSetDevice(g->ctxs[i]->device);
        imt_tree* t = nullptr;
        const imt_status st = tree_alloc(g->ctxs[i], n / g->world, true, &t);
        if (st != IMT_OK) {
            c0->last_error = g->ctxs[i]->last_error;
            imt_mtree_destroy(mt);
            return st;
        }
        mt->shards.push_back(t);
    }
    *out = mt;
    return IMT_OK;
}
imt_status mtree_rebuild_resident(imt_mtree* mt) { return mtree_rebuild(mt, nullptr); }
}  // namespace imt_host

// IndexedMerkleTree::new (src/utils.rs:20-57) fused with the leaf hashing (src/indexed_merkle_tree.rs:662-671) over all devices
extern "C" imt_status imt_multi_build_from_leaves(imt_multi* m, const void* preimages, size_t n, imt_mtree** out) {
    if (!m || !out) return IMT_ERR_INVALID_ARG;
    *out = nullptr;
    imt_group* g = &m->g;
    imt_ctx* c0 = g->ctxs[0];
    IMT_TRY(check_leaf_count(c0, n));
    if (!preimages) return fail(c0, IMT_ERR_INVALID_ARG, "null preimages");
    if (n < g->world || (n / g->world) * g->world != n) return fail(c0, IMT_ERR_INVALID_ARG, "fewer leaves than devices");
    imt_mtree* mt = nullptr;
    IMT_TRY(mtree_alloc(m, n, &mt));
    const imt_status st = mtree_rebuild(mt, preimages);
    if (st != IMT_OK) {
        imt_mtree_destroy(mt);
        return st;
    }
    *out = mt;
    return IMT_OK;
}
extern "C" imt_status imt_mtree_rebuild_from_leaves(imt_mtree* mt, const void* preimages) {
    if (!mt) return IMT_ERR_INVALID_ARG;
    if (!preimages) return fail(mt->m->g.ctxs[0], IMT_ERR_INVALID_ARG, "null preimages");
    return mtree_rebuild(mt, preimages);
}
extern "C" void imt_mtree_destroy(imt_mtree* mt) {
    if (!mt) return;
    for (imt_tree* t : mt->shards) imt_tree_destroy(t);
    delete mt;
}
extern "C" imt_tree* imt_mtree_shard(imt_mtree* mt, unsigned i) { return (mt && i < mt->shards.size()) ? mt->shards[i] : nullptr; }
extern "C" size_t imt_mtree_num_leaves(const imt_mtree* mt) { return mt ? mt->n_total : 0; }
extern "C" unsigned imt_mtree_depth(const imt_mtree* mt) { return (mt && !mt->shards.empty()) ? imt_tree_depth(mt->shards[0]) : 0; }
extern "C" imt_status imt_mtree_root(imt_mtree* mt, void* out_fe) {
    if (!mt || mt->shards.empty()) return IMT_ERR_INVALID_ARG;
    return imt_tree_root(mt->shards[0], out_fe);  // the cap is replicated: every shard reports the global root
}

namespace imt_host {
// the local shards of the group a tree belongs to, for the sharded calls of imt_indexed.cu
imt_status mtree_parts(imt_mtree* mt, imt_group** g, std::vector<imt_tree*>* trees) {
    if (!mt || mt->shards.empty()) return IMT_ERR_INVALID_ARG;
    *g = &mt->m->g;
    *trees = mt->shards;
    return IMT_OK;
}
}  // namespace imt_host
