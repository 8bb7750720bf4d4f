// Stream engine of the lane kernel (ros_lane.cu): one grid cell per LANE, one warp per block.
//
// Every phase of the integrator (rates, sparse sums, LU, triangular solves) is a STREAM of batches of four entries
// whose control words are the same for all 32 cells of the warp (kppgen/lsched.py emits them).  The per-cell data
// of an entry -- one double per lane, 256 contiguous bytes per warp in the [element][32 lanes] workspace -- is
// copied into a shared-memory ring by cp.async SIXTEEN batches (64 entries, 16 KB per warp) ahead of its use, so the
// two resident warps of an SM keep ~32 KB of HBM/L2 requests in flight.  (Loading into registers that far ahead
// does not work: a warp has six scoreboards, so a wait on one load waits for the youngest load sharing its
// scoreboard -- measured 1200 cycles per batch; cp.async groups complete in order and are waited for by count.)
// The control words are staged through a second small ring by 16-byte cp.async copies, one chunk of 16 records per
// refill, riding in the group of the chunk's first batch.  Everything a lane touches is private to the lane (its
// column of the workspace, of the ring and of the shared-memory vector), so the only warp-level synchronisation
// is the one that publishes a table chunk.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LANE_DEPTH 16            // batches in flight = records per table chunk
#define LANE_BATCH 4             // entries per batch
#define LANE_DRING_BYTES (LANE_DEPTH * LANE_BATCH * 256)
#define LANE_TRING_BYTES(RECQ) (3 * LANE_DEPTH * (RECQ) * 16)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned smem_dst, const void *gsrc)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A record is RECQ uint4 (RECQ = 2: four 32-bit entry words + a header; RECQ = 4: the same + four doubles).
// `tring` holds three chunks of 16 records, `dring` 64 entries x 32 lanes of doubles.  F supplies
//   void issue(const uint4 *rec, unsigned dst)      cp_async8 the batch's operands to dst + 256 * e (dst = this
//                                                   lane's column of the batch's ring slot)
//   void first(const uint4 *rec, const double *g)   before batch 0 is consumed (its ring operands have landed)
//   void consume(const uint4 *rec, const double *g, const uint4 *rec1, const double *g1)
//                                                   use batch t (g[32 * e] = entry e); rec1 / g1 are the record and
//                                                   the landed ring operands of batch t + 1, so that a consumer
//                                                   can load the shared-memory operands of t + 1 before it stores
// The table holds nchunk chunks followed by ONE chunk of zero records (so that the look-ahead never needs a
// bounds test); zero records neither load anything that matters nor compute.
// The batch loop is NOT unrolled: the code of all streams together must stay inside the instruction cache (an
// earlier version unrolled 16 batches for static ring addresses: 30 k instructions, and the one warp of an SM
// sub-partition spent half its cycles waiting for instruction fetches).
template <int RECQ, class F>
__device__ __forceinline__ void run_stream(const uint4 *__restrict__ tab, int nchunk, uint4 *tring, double *dring, F &f)
{
  constexpr int CHQ = LANE_DEPTH * RECQ;              // uint4 per chunk
  static_assert(CHQ % 32 == 0, "a chunk is copied by the 32 lanes in whole uint4s");
  constexpr int PER = CHQ / 32;
  constexpr int NREC = 3 * LANE_DEPTH;                // records in the table ring
  const int lane = threadIdx.x & 31;
  if (nchunk <= 0) return;
  const double *dl = dring + lane;
  const unsigned ds = (unsigned)__cvta_generic_to_shared(dring + lane);
  auto fill = [&](int c) {                            // c <= nchunk (the zero chunk)
#pragma unroll
    for (int r = 0; r < PER; r++) cp_async16(tring + (c % 3) * CHQ + r * 32 + lane, tab + (size_t)c * CHQ + r * 32 + lane);
  };
  __syncwarp();                                       // the previous stream's readers of the rings are done
  fill(0);
  fill(1);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();
#pragma unroll 1
  for (int u = 0; u < LANE_DEPTH; u++) {
    f.issue(tring + u * RECQ, ds + u * (LANE_BATCH * 256));
    cp_async_commit();
  }
  cp_async_wait<LANE_DEPTH - 1>();
  f.first(tring, dl);
  const int nbatch = nchunk * LANE_DEPTH;
  int ri = 0;                                         // index of batch t's record in the table ring
#pragma unroll 1
  for (int t = 0; t < nbatch; t++) {
    const int u = t & (LANE_DEPTH - 1);
    const int ri1 = (ri + 1 == NREC) ? 0 : ri + 1;
    int rii = ri + LANE_DEPTH;                        // record of batch t + 16
    if (rii >= NREC) rii -= NREC;
    cp_async_wait<LANE_DEPTH - 2>();                  // the groups of batches t and t + 1 have landed
    if (u == 0) __syncwarp();                         // ... including (all lanes' parts of) the next table chunk
    f.consume(tring + ri * RECQ, dl + u * (LANE_BATCH * 32), tring + ri1 * RECQ, dl + ((u + 1) & (LANE_DEPTH - 1)) * (LANE_BATCH * 32));
    if (u == 0 && (t >> 4) + 2 <= nchunk) fill((t >> 4) + 2);   // into the slot every lane left before the __syncwarp above
    f.issue(tring + rii * RECQ, ds + u * (LANE_BATCH * 256));
    cp_async_commit();
    ri = ri1;
  }
  cp_async_wait<0>();
}
