/* paraiso_b200.h — the C ABI between a Paraiso host class and the B200 kernels.
 *
 * What it replaces.  The reference's CUDA flavour emits one translation unit in which the member function
 * of every subkernel launches its `__global__` helper directly:
 *     Language/Paraiso/Generator/PlanTrans.hs:295-314   (the `<sub>_inner` kernel: flat grid-stride loop)
 *     Language/Paraiso/Generator/PlanTrans.hs:318-343   (`<sub>_inner<<<grid, block>>>(raw pointers...)`)
 *     Language/Paraiso/Generator/PlanTrans.hs:225-258   (kernel driver: subkernel calls, then `static = manifest` copies)
 *     Language/Paraiso/Generator/PlanTrans.hs:719-720   (om_reduce_* via Thrust, result returned to the host)
 * There is no FFI in the reference.  This header is the boundary a maintainer binds instead: one
 * `extern "C"` launcher per fused GPU stage of every OM kernel, plain pointers and ints only.
 *
 * Per machine `<Name>` the generator emits `<Name>_abi.h` (copied here as om_<Name>_abi.h) declaring
 *     int om_<Name>_abi_version(void);
 *     int om_<Name>_<kernel>_stage<k>(const OmGeomC* g, void* const* cur, void* const* alt,
 *                                     void* sc, void* scratch, void* stream);
 *     int om_<Name>_<kernel>_stage<k>_occupancy(void);
 *     int om_<Name>_<kernel>_scalars(const OmGeomC* g, void* sc, void* stream);
 *
 * Calling convention
 *   g        geometry of this rank's slab (sizes are run-time values; margins / boundary kinds are baked in)
 *   cur/alt  arrays of device pointers indexed by static-variable index; a stage reads `cur[i]` and writes
 *            `alt[i]`; after all stages of an OM kernel the host swaps cur/alt of every stored array
 *            (the reference's store-after-compute semantics, PlanTrans.hs:228, without the copy)
 *   sc       device scalar slots, 8 bytes each: slot i < nstatics is static variable i (Scalar realm),
 *            the following slots hold Reduce results
 *   scratch  >= 256 + 8 * CTAs * reduces bytes of device memory, zeroed once at creation
 *   stream   cudaStream_t
 *   return   0, or the cudaError_t of the failed launch; nothing throws across the boundary
 *
 * Memory contract: every array is `rows` x `pitch` elements, row-major, axis 0 fastest
 * (PlanTrans.hs:449-454), interior cell (0, y0) at (row yorg, column xorg), ghost / margin cells around
 * it, and OM_APRON_ROWS zero-initialised rows allocated above row 0 and below row rows-1 (kernels do not
 * bounds-check their pipeline fill).  `pitch` is a multiple of 32 elements, `xorg` a multiple of 32.
 * On Cyclic axes the ghost cells of an array must be valid before a stage reads it; stages keep them valid
 * for the arrays they write (x always; y when wrap_y_local), the host refreshes them after host writes and,
 * with several ranks, exchanges the y ghost rows.
 * Rank-3 machines (ABI version 2) stack `nz + gz_lo + gz_hi` planes of `rows * pitch` elements along axis 2
 * (`plane` = that stride, interior plane 0 at device plane `zorg`; the aprons surround the whole stack); a launch
 * covers device planes [own_z0, own_z1) with one layer of CTAs per plane, and the host copies the ghost planes of a
 * Cyclic axis 2 after each store; with several ranks the slab is cut along axis 2 (`z0` = global index of local plane 0,
 * `nzl` local planes).  Rank-1 / rank-2 callers pass nz = 1, plane = 0, own_z0 = 0, own_z1 = 1, z0 = 0, nzl = 1.
 *
 * Several ranks (ABI version 3): `g->bfirst = 1` launches a stage in boundary-first chunk order — blockIdx.y 0 computes the
 * LAST chunk of rows, blockIdx.y c > 0 chunk c - 1 — so the rows the slab's neighbours read (`sig_lo`: rows of the first
 * chunk, `sig_hi`: rows of the last chunk) are written by the first wave of CTAs.  The last of those CTAs raises a flag in
 * the scratch header, and
 *     int om_<Name>_wait_boundary(void* scratch, void* stream);
 * enqueues a one-thread kernel on `stream` (the host's high-priority communication stream) that completes when the flag is
 * up: ncclSend / ncclRecv of the ghost rows enqueued behind it overlap the remaining waves of the same launch.  This replaces
 * separate boundary launches; the reference has no counterpart (single device, PlanTrans.hs:318-343).  Single-rank callers
 * pass bfirst = sig_lo = sig_hi = 0.
 */
#pragma once
#include "om_Life_abi.h"
#include "om_Hydro_abi.h"
