/*
 * slrbs_pile.h -- C ABI of libslrbs_pile.so: the inter-GPU exchange of one large pile cut into slabs (SURVEY.md 8(e),
 * "one large scene": neighbour ncclSend / ncclRecv of the halo bodies' state). The reference is single-process and has
 * no counterpart; what this replaces is the per-step halo refresh that RigidBodySystem::step would need between the
 * integration of one step (src/rigidbody/RigidBodySystem.cpp:99-120) and the detection of the next (:77-81) when the
 * bodies of one scene live on several GPUs.
 *
 * Everything is issued on the simulation context's own stream (slrbs_stream): pack kernel -> grouped point-to-point with
 * the left and right slab neighbour -> unpack kernel, so a step needs no host synchronisation. One process per GPU;
 * NCCL over NVLink / NVSwitch.
 */
#ifndef SLRBS_PILE_H
#define SLRBS_PILE_H

#include <stddef.h>

#include "slrbs_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct slrbs_pile_comm slrbs_pile_comm;

const char* slrbs_pile_comm_last_error(void);
/* ncclGetUniqueId: 128 bytes, made on one rank and handed to the others by the caller (file, socket, MPI, torch ...) */
int  slrbs_pile_comm_unique_id(void* out128);
/* ncclCommInitRank on cuda:device; slabs are ordered by rank (rank - 1 is the left neighbour) */
int  slrbs_pile_comm_create(slrbs_pile_comm** out, int rank, int world, int device, const char* unique_id128);
void slrbs_pile_comm_destroy(slrbs_pile_comm* comm);
/* One halo refresh on ctx's stream, asynchronous: slrbs_halo_pack(ctx, send) -> the first n_to_left states of `send` go to
 * rank - 1 and the next n_to_right to rank + 1, `recv` receives n_from_left states from rank - 1 followed by n_from_right
 * from rank + 1 -> slrbs_halo_unpack(ctx, recv). A state is 13 floats (x, q, xdot, omega). Buffers are device memory. */
int  slrbs_pile_comm_halo_exchange(slrbs_pile_comm* comm, slrbs_ctx* ctx, void* send, void* recv,
                                   int n_to_left, int n_to_right, int n_from_left, int n_from_right);
/* sum of a float table over all ranks, in place, on ctx's stream (the exported owner states of a re-partition) */
int  slrbs_pile_comm_allreduce_sum(slrbs_pile_comm* comm, slrbs_ctx* ctx, float* table, size_t n);

#ifdef __cplusplus
}
#endif
#endif
