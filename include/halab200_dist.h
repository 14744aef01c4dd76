/* halab200_dist.h — multi-GPU extension of the C ABI (one process per GPU, NCCL over NVLink/NVSwitch).
 *
 * The reference has no multi-device path at all (one gpu_engine == one device id, gpu/hala_gpu_engine.hpp:55,60; its tests
 * only loop over devices, tests/solvers_tests.cpp:58-68), so nothing here replaces a reference interface: these entry
 * points are what BASELINE.json's north_star adds — 1-D row-block partition, ghost-entry halo exchange, scalar all-reduce.
 * The ghost maps themselves are built by the host layer (hala_b200/partition.py) and handed over with hb_dist_set_plan.
 * NCCL is resolved at run time (dlopen of libnccl.so.2), so libhalab200.so has no hard dependency on it.
 */
#ifndef HALAB200_DIST_H
#define HALAB200_DIST_H
#include "halab200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_dist hb_dist;
enum { HB_NCCL_ID_BYTES = 128 };

/* rank 0 creates the id, the launcher broadcasts the 128 bytes (torch.distributed / MPI / a file), every rank calls create */
int hb_dist_unique_id(void *id128);
int hb_dist_create(hb_ctx *ctx, int rank, int world, const void *id128, hb_dist **dist);
int hb_dist_destroy(hb_dist *dist);
int hb_dist_info(const hb_dist *dist, int *rank, int *world);
/* which transport the last hb_dist_cg on the current plan used: peer memory (every rank maps every other rank's exchange
 * buffer through CUDA IPC; halo entries and scalar partials are stored straight into the peers' memory over NVLink by the
 * iteration kernels, no collective call per iteration) or NCCL (send/recv + all-reduce).  Peer is chosen when all ranks can
 * map each other (one NVLink domain, <= 16 ranks, <= 8 halo neighbours); HB_DIST_PEER=0 in the environment forces NCCL. */
enum { HB_TRANSPORT_NCCL = 0, HB_TRANSPORT_PEER = 1 };
int hb_dist_transport(const hb_dist *dist, int *transport);

/* Exchange plan of this rank.  n_owned rows/columns are owned; ghost columns are numbered n_owned .. n_owned + n_ghost - 1
 * in the order they are received: neighbour 0's block first, then neighbour 1's, ...
 *   neigh[k]            rank of neighbour k (0 <= k < nneigh), ascending
 *   send_count[k]       how many owned entries neighbour k needs;  send_idx (DEVICE, concatenated) = their local indices
 *   recv_count[k]       how many ghost entries come from neighbour k (sum == n_ghost)                                        */
int hb_dist_set_plan(hb_dist *dist, int n_owned, int n_ghost, int nneigh, const int *neigh,
                     const int *send_count, const int *recv_count, const int *send_idx_dev);

/* x_ext = [x_owned | ghosts]: the requested owned entries travel to the neighbours, ghosts land in x_ext + n_owned.  Over peer memory
 * once that transport is up for the element size (two small kernels: push into the neighbours' exchange buffers + flags, then
 * wait + copy), otherwise packed and sent with grouped ncclSend/ncclRecv (the *_nccl form always takes that route). */
int hb_dist_halo_exchange(hb_dist *dist, int dtype, void *x_ext);
int hb_dist_halo_exchange_nccl(hb_dist *dist, int dtype, void *x_ext);
/* in-place sum over ranks of `count` scalars of `dtype` that live on the device: one block exchanging through the peers' mailboxes
 * (count <= 66, summed in rank order: identical bits on all ranks), or ncclAllReduce */
int hb_dist_allreduce_sum(hb_dist *dist, int dtype, void *dev_scalars, int count);
int hb_dist_allreduce_sum_nccl(hb_dist *dist, int dtype, void *dev_scalars, int count);
/* collective: brings the peer transport up for `dtype` when all ranks can map each other (hb_dist_cg / hb_dist_gmres call it) */
int hb_dist_prepare_transport(hb_dist *dist, int dtype);
/* collective, closes a sequence of hb_dist_halo_exchange / hb_dist_allreduce_sum calls that ran over peer memory: *timed_out = 1 on ALL
 * ranks when a wait on a peer's flag gave up on any rank (the sums were poisoned with NaN); the communicator then drops to NCCL for good
 * and the caller redoes its work.  hb_dist_cg / hb_dist_gmres do this themselves. */
int hb_dist_finish_transport(hb_dist *dist, int *timed_out);
/* diagnostics: the peer protocol's sequence numbers (equal on all ranks after every solve), how many solves fell back from peer memory
 * to NCCL after a time-out, and how many times the start-of-solve agreement found the ranks' sequence numbers different */
int hb_dist_debug_info(const hb_dist *dist, unsigned long long *epoch, unsigned long long *vepoch, int *peer_fallbacks, int *epoch_repairs);

/* Row-partitioned CG.  csr = local rows with columns renumbered to [owned | ghosts] (cols == n_owned + n_ghost);
 * b, x = owned parts.  Same recurrence, counter and stop test as hb_cg; per iteration: halo exchange of p, SpMV fused with
 * the local <p,Ap>, sum over ranks, fused update + local ||r||^2, sum over ranks, x and direction update.  All ranks
 * return the same iteration count and residual (the sums are formed in rank order on every rank).                                                                                          */
int hb_dist_cg(hb_dist *dist, const hb_csr *csr, const void *b, void *x, double tol, int max_iter, int *iters, double *res);
/* Row-partitioned GMRES(m): hb_gmres with the halo of each basis vector exchanged in place before its SpMV, the k Gram-Schmidt
 * coefficients and the norm all-reduced (k + 1 scalars per inner iteration), Givens/Hessenberg replicated on every host. */
int hb_dist_gmres(hb_dist *dist, const hb_csr *csr, const void *b, void *x, double tol, int max_outer, int restart, int cproj,
                  int *iters, double *res);
/* y_owned = A_local * [x_owned | ghosts(x)]  (one halo exchange + one SpMV); x_ext must have room for the ghosts */
int hb_dist_spmv(hb_dist *dist, const hb_csr *csr, void *x_ext, void *y);

#ifdef __cplusplus
}
#endif
#endif
