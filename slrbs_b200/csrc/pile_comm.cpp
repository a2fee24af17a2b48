// pile_comm.cpp -- libslrbs_pile.so (include/slrbs_pile.h): NCCL point-to-point halo exchange between slab neighbours,
// ordered on the simulation context's stream. Kept out of libslrbs_b200.so so that the batched-scene path (no collective,
// SURVEY.md 8(e)) carries no NCCL dependency.
#include "../../include/slrbs_pile.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstring>
#include <string>

namespace {
thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return -1; }
enum { STATE_W = 13 };
}

struct slrbs_pile_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

#define NC(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) return fail(std::string(#call) + ": " + ncclGetErrorString(r_));    \
    } while (0)

extern "C" {

const char* slrbs_pile_comm_last_error(void) { return g_err.c_str(); }

int slrbs_pile_comm_unique_id(void* out128)
{
    static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
    if (!out128) return fail("slrbs_pile_comm_unique_id: null buffer");
    ncclUniqueId id;
    NC(ncclGetUniqueId(&id));
    std::memcpy(out128, &id, sizeof id);
    return 0;
}

int slrbs_pile_comm_create(slrbs_pile_comm** out, int rank, int world, int device, const char* unique_id128)
{
    if (!out || !unique_id128 || world < 1 || rank < 0 || rank >= world) return fail("slrbs_pile_comm_create: bad argument");
    *out = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return fail("slrbs_pile_comm_create: cudaSetDevice failed (no CPU fallback)");
    ncclUniqueId id;
    std::memcpy(&id, unique_id128, sizeof id);
    slrbs_pile_comm* c = new slrbs_pile_comm();
    c->rank = rank; c->world = world; c->device = device;
    const ncclResult_t r = ncclCommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) { delete c; return fail(std::string("ncclCommInitRank: ") + ncclGetErrorString(r)); }
    *out = c;
    return 0;
}

void slrbs_pile_comm_destroy(slrbs_pile_comm* c)
{
    if (!c) return;
    if (c->comm) ncclCommDestroy(c->comm);
    delete c;
}

int slrbs_pile_comm_halo_exchange(slrbs_pile_comm* c, slrbs_ctx* ctx, void* send, void* recv, int n_to_left, int n_to_right,
                                  int n_from_left, int n_from_right)
{
    if (!c || !ctx || n_to_left < 0 || n_to_right < 0 || n_from_left < 0 || n_from_right < 0)
        return fail("slrbs_pile_comm_halo_exchange: bad argument");
    if ((c->rank == 0 && (n_to_left || n_from_left)) || (c->rank == c->world - 1 && (n_to_right || n_from_right)))
        return fail("slrbs_pile_comm_halo_exchange: the first / last slab has no left / right neighbour");
    cudaStream_t s = (cudaStream_t)slrbs_stream(ctx);
    if (slrbs_halo_pack(ctx, send) < 0) return fail(std::string("slrbs_halo_pack: ") + slrbs_last_error());
    float* sb = (float*)send;
    float* rb = (float*)recv;
    NC(ncclGroupStart());
    if (n_to_left)    NC(ncclSend(sb, (size_t)n_to_left * STATE_W, ncclFloat, c->rank - 1, c->comm, s));
    if (n_to_right)   NC(ncclSend(sb + (size_t)n_to_left * STATE_W, (size_t)n_to_right * STATE_W, ncclFloat, c->rank + 1, c->comm, s));
    if (n_from_left)  NC(ncclRecv(rb, (size_t)n_from_left * STATE_W, ncclFloat, c->rank - 1, c->comm, s));
    if (n_from_right) NC(ncclRecv(rb + (size_t)n_from_left * STATE_W, (size_t)n_from_right * STATE_W, ncclFloat, c->rank + 1, c->comm, s));
    NC(ncclGroupEnd());
    if (slrbs_halo_unpack(ctx, recv) < 0) return fail(std::string("slrbs_halo_unpack: ") + slrbs_last_error());
    return 0;
}

int slrbs_pile_comm_allreduce_sum(slrbs_pile_comm* c, slrbs_ctx* ctx, float* table, size_t n)
{
    if (!c || !ctx || !table) return fail("slrbs_pile_comm_allreduce_sum: bad argument");
    NC(ncclAllReduce(table, table, n, ncclFloat, ncclSum, c->comm, (cudaStream_t)slrbs_stream(ctx)));
    return 0;
}

}  // extern "C"
