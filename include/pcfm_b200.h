/*
 * pcfm_b200.h — C ABI of libpcfm_b200.so: B200 (sm_100a) kernels for the point-cloud hot path of
 * ruanjinchen/Point-Cloud-Flow-Matching (PVCNN point-voxel primitives + Chamfer-3D + approximate EMD).
 *
 * Boundary contract
 *   - extern "C", plain pointers and sizes, no torch/ATen types.  Every pointer is a DEVICE pointer unless its
 *     name ends in `_host`.  All tensors are dense, contiguous, fp32 / int32, laid out exactly as the
 *     reference's tensors (PVCNN ops: channel-first [B,C,N] / [B,3,N]; Chamfer/EMD: [B,N,3]).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Every entry point is
 *     asynchronous, never synchronises the host, and launches only on `stream` (the reference launches
 *     K1-K5, K11, K15-K20 on the legacy stream) so callers may CUDA-graph capture.
 *   - Return value: 0 on success, otherwise a cudaError_t (positive) or PCFM_E_* (negative).
 *   - Outputs are FULLY overwritten (no pre-zeroing needed) unless the comment says "accumulates".
 *   - `workspace` is caller-owned scratch of at least pcfm_*_workspace(...) bytes, 256-byte aligned; it may
 *     be reused by the next call on the same stream.
 *   - Index arithmetic is 64-bit inside the kernels (the reference overflows int at B*C*r^3 >= 2^31).
 *
 * Each entry point cites the reference interface (file:line under /root/reference) it replaces.
 */
#ifndef PCFM_B200_H
#define PCFM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCFM_ABI_VERSION 1

#define PCFM_E_BADARG   (-1)  /* NULL pointer / negative size / unsupported size */
#define PCFM_E_WORKSPACE (-2) /* workspace too small */

typedef void* pcfm_stream_t; /* cudaStream_t */

/* Library / build info. */
int pcfm_abi_version(void);
/* Returns a static string such as "pcfm_b200 sm_100a nvcc 12.9". */
const char* pcfm_build_info(void);
/* Number of kernel launches issued by this library in the calling process since load (or last reset). */
int64_t pcfm_launch_count(void);
void pcfm_reset_launch_count(void);
/* Identity of the CUDA-graph capture that `stream` is recording, 0 when it is not capturing
 * (cudaStreamGetCaptureInfo).  Host-side caches (pcfm.nn.Voxelization's per-coords memo) key on it: a tensor computed
 * outside a capture must not be reused inside one — the graph would replay without the kernels that produce it. */
unsigned long long pcfm_stream_capture_id(pcfm_stream_t stream);

/* Tuning knob: "variant" selects alternative kernel implementations for A/B measurement.
 * key: "devox_bwd" (generic inds/wgts entry: 0 = sorted-segment gather [default], 1 = global RED atomics),
 *      "vox_fwd"   (0 = cell plan [default], 1 = global RED atomics, 2 = generic sorted-segment gather),
 *      "stage"     (0 = TMA bulk staging where aligned [default], 1 = plain ld/st staging),
 *      "devox_fwd" (host-side choice read by pcfm/backend.py: 0 = cell-plan walk where supported [default],
 *                   1 = plane-staging point-lane kernels). */
int pcfm_set_option(const char* key, int value);

/* ------------------------------------------------------------------------------------------------
 * Voxelisation (third_party/pvcnn/modules/functional/src/voxelization/vox.cpp:17-43,54-76;
 *               kernels vox.cu:18-34,48-72,86-110)
 * ---------------------------------------------------------------------------------------------- */

/* Voxelization.forward tail (third_party/pvcnn/modules/voxelization.py:19-24), one fused kernel instead of six
 * elementwise torch kernels.  centred f32[B,3,N] = coords - mean over N (voxelization.py:18, computed by the
 * caller so that the reduction order stays the caller's); scale f32[B] = 2*max||centred|| + eps, or NULL for
 * normalize=False.  Per element, each step rounded to fp32 exactly like the torch ops it replaces:
 *   normalize:  v = centred / scale + 0.5        else:  v = (centred + 1) * 0.5
 *   v = clamp(v * r, 0, r-1)  -> norm_coords f32[B,3,N];   vox_coords i32[B,3,N] = rint(v) (half to even). */
int pcfm_voxel_coords(const float* centred, const float* scale, int b, int n, int r, float* norm_coords,
                      int* vox_coords, pcfm_stream_t stream);

/* avg_voxelize_forward (vox.cpp:17-43): features f32[B,C,N], coords i32[B,3,N] (already voxel integers in
 * [0,r-1]) -> out f32[B,C,r^3], ind i32[B,N] (= x*r^2+y*r+z), cnt i32[B,r^3].
 * out[b,c,v] = sum over points with ind==v of features*(1/cnt[v]); empty voxels are 0.
 * The sum runs in ascending point order (one legal ordering of the reference's atomics) => deterministic.
 * Limits of the cell plan (24-bit fields): r <= 256 and N below ~16.7 million points per cloud; beyond them this entry,
 * the _dual / _planned entries and the trilinear *_coords / *_planned / *_cells entries return PCFM_E_BADARG. */
size_t pcfm_avg_voxelize_forward_workspace(int b, int c, int n, int r);
int pcfm_avg_voxelize_forward(const float* features, const int* coords, int b, int c, int n, int r,
                              float* out, int* ind, int* cnt,
                              void* workspace, size_t workspace_bytes, pcfm_stream_t stream);

/* avg_voxelize_backward (vox.cpp:54-76): grad_y f32[B,C,S], ind i32[B,N], cnt i32[B,S] -> grad_x f32[B,C,N]
 * grad_x[b,c,i] = grad_y[b,c,ind[i]] * (1/cnt[ind[i]]). */
size_t pcfm_avg_voxelize_backward_workspace(int b, int c, int n, int s);
int pcfm_avg_voxelize_backward(const float* grad_y, const int* ind, const int* cnt, int b, int c, int n, int s,
                               float* grad_x, void* workspace, size_t workspace_bytes, pcfm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Trilinear devoxelisation (src/interpolate/trilinear_devox.cpp:18-55,67-91; kernels
 *                           trilinear_devox.cu:21-105,119-162)
 * ---------------------------------------------------------------------------------------------- */

/* trilinear_devoxelize_forward (trilinear_devox.cpp:18-55): coords f32[B,3,N] in [0,r-1],
 * features f32[B,C,r^3] -> outs f32[B,C,N]; if is_training also inds i32[B,8,N], wgts f32[B,8,N]
 * (corner order 000,001,010,011,100,101,110,111; hi offsets only where frac>0, trilinear_devox.cu:64-75).
 * inds/wgts may be NULL when is_training==0. */
int pcfm_trilinear_devoxelize_forward(const float* coords, const float* features, int b, int c, int n, int r,
                                      int is_training, float* outs, int* inds, float* wgts,
                                      pcfm_stream_t stream);

/* trilinear_devoxelize_backward (trilinear_devox.cpp:67-91): grad_y f32[B,C,N], inds i32[B,8,N],
 * wgts f32[B,8,N] -> grad_x f32[B,C,r^3];  grad_x[b,c,inds[b,k,i]] += wgts[b,k,i]*grad_y[b,c,i].
 * Works for arbitrary inds in [0,r^3).  Deterministic (sorted segmented sums) in the default variant. */
size_t pcfm_trilinear_devoxelize_backward_workspace(int b, int c, int n, int r);
int pcfm_trilinear_devoxelize_backward(const float* grad_y, const int* inds, const float* wgts,
                                       int b, int c, int n, int r, float* grad_x,
                                       void* workspace, size_t workspace_bytes, pcfm_stream_t stream);

/* Same result as pcfm_trilinear_devoxelize_backward for the inds/wgts that pcfm_trilinear_devoxelize_forward derives
 * from `coords` (f32[B,3,N] in [0,r-1]), but takes the coordinates themselves: the 8 weights are recomputed with the
 * forward's exact arithmetic and every point is visited once per channel (cell-sorted plan).  This is the entry the
 * autograd wrapper uses (it saves `coords` next to the reference's inds/wgts).  r <= 37: the channel plane accumulates
 * in shared memory (no global atomics, no memset); 38 <= r <= 256: grad_x is zeroed and the per-cell corner sums are
 * added with global atomics (8 per cell instead of the reference's 8 per point). */
size_t pcfm_trilinear_devoxelize_backward_coords_workspace(int b, int c, int n, int r);
int pcfm_trilinear_devoxelize_backward_coords(const float* grad_y, const float* coords, int b, int c, int n, int r,
                                              float* grad_x, void* workspace, size_t workspace_bytes,
                                              pcfm_stream_t stream);

/* Plan reuse.  pcfm_avg_voxelize_forward and pcfm_trilinear_devoxelize_backward_coords group the points by voxel
 * cell first (the "cell plan", a pure function of coords, b, n, r) and leave that plan in the caller's workspace.  The
 * PVConv blocks of one stage of the reference's ContextNet see the SAME coordinates (models.py:385-389 passes them
 * through, pvconv.py:39 returns them unchanged), so the second block - and the second backward - can skip the build:
 * the two entries below take the workspace of an earlier call with identical (coords, b, n, r), untouched since, and
 * launch only the row kernel.  Results are those of the unplanned entries on the same coords.  The caller owns the
 * validity of the plan (pcfm/backend.py keys its memo on tensor identity + version + CUDA-graph capture + stream). */
/* pcfm_avg_voxelize_forward that ALSO leaves the trilinear plan of `float_coords` (f32[B,3,N] in [0,r-1], the
 * coordinates the PVConv block will devoxelize at, pvconv.py:35-37) in `tri_workspace`
 * (pcfm_trilinear_devoxelize_backward_coords_workspace bytes), ready for pcfm_trilinear_devoxelize_backward_planned in
 * the backward pass.  Both plans are built by one kernel: a single build is latency-bound on 8 CTAs per cloud, the
 * second one runs beside it on idle SMs.  `workspace` afterwards holds the mean plan as after
 * pcfm_avg_voxelize_forward. */
int pcfm_avg_voxelize_forward_dual(const float* features, const int* coords, const float* float_coords, int b, int c,
                                   int n, int r, float* out, int* ind, int* cnt, void* workspace,
                                   size_t workspace_bytes, void* tri_workspace, size_t tri_workspace_bytes,
                                   pcfm_stream_t stream);
int pcfm_avg_voxelize_forward_planned(const float* features, int b, int c, int n, int r, float* out,
                                      const void* plan_workspace, size_t plan_workspace_bytes,
                                      pcfm_stream_t stream);
int pcfm_trilinear_devoxelize_backward_planned(const float* grad_y, int b, int c, int n, int r, float* grad_x,
                                               const void* plan_workspace, size_t plan_workspace_bytes,
                                               pcfm_stream_t stream);

/* trilinear_devoxelize_forward (trilinear_devox.cpp:18-55) as a walk over the trilinear cell plan: same outputs as
 * pcfm_trilinear_devoxelize_forward (outs within rounding: the 8 products are summed along z, y, x instead of in
 * corner order; inds / wgts bit-identical).  The points are visited cell by cell, so the 8 corner values of a cell are
 * read once per cell and channel, and the output row leaves shared memory with one bulk store.
 * `workspace`: pcfm_trilinear_devoxelize_backward_coords_workspace(b, c, n, r) bytes.  planned == 0: the plan of
 * `coords` is built into it first (and stays there for pcfm_trilinear_devoxelize_backward_planned and later calls
 * with planned == 1); planned != 0: it already holds the plan of the same (coords, b, n, r) — left by this entry,
 * pcfm_avg_voxelize_forward_dual or pcfm_trilinear_devoxelize_backward_coords — and `coords` is read only for
 * inds / wgts (may be NULL when is_training == 0).
 * pcfm_trilinear_devoxelize_forward_cells_ok: 1 when this entry supports the shape (the output row of N floats must
 * fit in shared memory: N <= 57k; 24-bit plan fields), else 0 — use pcfm_trilinear_devoxelize_forward. */
int pcfm_trilinear_devoxelize_forward_cells_ok(int b, int c, int n, int r);
int pcfm_trilinear_devoxelize_forward_cells(const float* coords, const float* features, int b, int c, int n, int r,
                                            int is_training, float* outs, int* inds, float* wgts, void* workspace,
                                            size_t workspace_bytes, int planned, pcfm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Ball query / grouping / gather / FPS / 3-NN  (src/ball_query, src/grouping, src/sampling,
 *                                               src/interpolate/neighbor_interpolate.*)
 * ---------------------------------------------------------------------------------------------- */

/* ball_query_forward (ball_query.cpp:6-30, ball_query.cu:19-50): centers f32[B,3,M], points f32[B,3,N]
 * -> idx i32[B,M,U]: first U point indices (ascending) with d2 < radius^2 (strict, d2 evaluated as in the
 * reference SASS: fma(dz,dz,fma(dx,dx,dy*dy)) on centre-point); remaining slots = first hit; no hit -> 0. */
int pcfm_ball_query(const float* centers, const float* points, int b, int n, int m, float radius, int u,
                    int* idx, pcfm_stream_t stream);

/* grouping_forward (grouping.cpp:6-24): features f32[B,C,N], indices i32[B,M,U] -> out f32[B,C,M,U]. */
int pcfm_grouping_forward(const float* features, const int* indices, int b, int c, int n, int m, int u,
                          float* out, pcfm_stream_t stream);
/* grouping_backward (grouping.cpp:26-44): grad_y f32[B,C,M,U], indices -> grad_x f32[B,C,N] (scatter-add). */
size_t pcfm_grouping_backward_workspace(int b, int c, int n, int m, int u);
int pcfm_grouping_backward(const float* grad_y, const int* indices, int b, int c, int n, int m, int u,
                           float* grad_x, void* workspace, size_t workspace_bytes, pcfm_stream_t stream);

/* gather_features_forward (sampling.cpp:6-24): features f32[B,C,N], indices i32[B,M] -> out f32[B,C,M]. */
int pcfm_gather_features_forward(const float* features, const int* indices, int b, int c, int n, int m,
                                 float* out, pcfm_stream_t stream);
/* gather_features_backward (sampling.cpp:26-41): grad_y f32[B,C,M], indices -> grad_x f32[B,C,N]. */
size_t pcfm_gather_features_backward_workspace(int b, int c, int n, int m);
int pcfm_gather_features_backward(const float* grad_y, const int* indices, int b, int c, int n, int m,
                                  float* grad_x, void* workspace, size_t workspace_bytes, pcfm_stream_t stream);

/* furthest_point_sampling_forward (sampling.cpp:43-58, sampling.cu:86-167): coords f32[B,3,N] ->
 * indices i32[B,M].  indices[0]=0; round j picks arg-max of the running min squared distance with the
 * reference's tie-break (ties -> smallest k mod 512, then smallest k).  workspace: B*N floats. */
size_t pcfm_furthest_point_sampling_workspace(int b, int n, int m);
int pcfm_furthest_point_sampling(const float* coords, int b, int n, int m, int* indices,
                                 void* workspace, size_t workspace_bytes, pcfm_stream_t stream);

/* three_nearest_neighbors_interpolate_forward (neighbor_interpolate.cpp:6-40, .cu:20-116):
 * points f32[B,3,N], centers f32[B,3,M], centers_features f32[B,C,M] ->
 * out f32[B,C,N], idx i32[B,3,N], wgt f32[B,3,N]. */
int pcfm_three_nn_interpolate_forward(const float* points, const float* centers, const float* centers_features,
                                      int b, int c, int n, int m, float* out, int* idx, float* wgt,
                                      pcfm_stream_t stream);
/* three_nearest_neighbors_interpolate_backward (neighbor_interpolate.cpp:42-65, .cu:145-170):
 * grad_y f32[B,C,N], idx i32[B,3,N], wgt f32[B,3,N] -> grad_x f32[B,C,M]. */
size_t pcfm_three_nn_interpolate_backward_workspace(int b, int c, int n, int m);
int pcfm_three_nn_interpolate_backward(const float* grad_y, const int* idx, const float* wgt,
                                       int b, int c, int n, int m, float* grad_x,
                                       void* workspace, size_t workspace_bytes, pcfm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Chamfer 3D (third_party/ChamferDistancePytorch/chamfer3D/chamfer_cuda.cpp:17-26,
 *             chamfer3D.cu:12-154 forward, :155-195 backward)
 * ---------------------------------------------------------------------------------------------- */

/* chamfer_3D.forward (chamfer_cuda.cpp:17-19): xyz1 f32[B,N,3], xyz2 f32[B,M,3] ->
 * dist1 f32[B,N] (min_j squared distance), idx1 i32[B,N] (lowest-index arg-min), dist2 f32[B,M], idx2 i32[B,M].
 * Distances are evaluated exactly as the reference SASS: d = fma(dz,dz,fma(dx,dx,dy*dy)) on (other - self). */
int pcfm_chamfer3d_forward(const float* xyz1, const float* xyz2, int b, int n, int m,
                           float* dist1, float* dist2, int* idx1, int* idx2, pcfm_stream_t stream);
/* Same outputs, bit for bit (distances, and arg-min indices including the lowest-index tie rule), with every squared
 * distance evaluated ONCE for both directions: row minima in registers, column minima combined through 64-bit
 * (distance bits << 32 | index) atomicMin keys in `workspace` (pcfm_chamfer3d_forward_workspace bytes).  This is the
 * entry pcfm.backend.chamfer.forward uses; the entry above (two directional passes, no workspace) is kept for callers
 * without scratch memory. */
size_t pcfm_chamfer3d_forward_workspace(int b, int n, int m);
int pcfm_chamfer3d_forward_ws(const float* xyz1, const float* xyz2, int b, int n, int m,
                              float* dist1, float* dist2, int* idx1, int* idx2,
                              void* workspace, size_t workspace_bytes, pcfm_stream_t stream);
/* chamfer_3D.backward (chamfer_cuda.cpp:22-26): ACCUMULATES into gradxyz1 f32[B,N,3], gradxyz2 f32[B,M,3]
 * (caller zeroes them, as dist_chamfer_3D.py:75-79 does). */
int pcfm_chamfer3d_backward(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                            const int* idx1, const int* idx2, int b, int n, int m,
                            float* gradxyz1, float* gradxyz2, pcfm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Approximate EMD (third_party/PyTorchEMD/cuda/emd.cpp:23-27, emd_kernel.cu:24-396)
 * ---------------------------------------------------------------------------------------------- */

/* approxmatch_forward (emd_kernel.cu:169-191): xyz1 f32[B,N,3], xyz2 f32[B,M,3] -> match f32[B,M,N].
 * workspace: B*2*(N+M) floats (the reference's `temp`). */
size_t pcfm_emd_approxmatch_workspace(int b, int n, int m);
int pcfm_emd_approxmatch_forward(const float* xyz1, const float* xyz2, int b, int n, int m, float* match,
                                 void* workspace, size_t workspace_bytes, pcfm_stream_t stream);
/* matchcost_forward (emd_kernel.cu:255-277): -> cost f32[B] = sum_{k,l} |x1_k-x2_l|^2 * match[l,k]. */
size_t pcfm_emd_matchcost_workspace(int b, int n, int m);
int pcfm_emd_matchcost_forward(const float* xyz1, const float* xyz2, const float* match, int b, int n, int m,
                               float* cost, void* workspace, size_t workspace_bytes, pcfm_stream_t stream);
/* matchcost_backward (emd_kernel.cu:371-396): grad_cost f32[B] -> grad1 f32[B,N,3], grad2 f32[B,M,3]. */
int pcfm_emd_matchcost_backward(const float* grad_cost, const float* xyz1, const float* xyz2, const float* match,
                                int b, int n, int m, float* grad1, float* grad2, pcfm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * All-pairs evaluation blocks (new capability used by the sharded MMD/COV/1-NNA harness; each entry is
 * defined as the reference operator applied to the pair (S[i], R[j]) — SURVEY.md section 8(e)).
 * ---------------------------------------------------------------------------------------------- */

/* Chamfer distance for D-dimensional points, D in {2, 5, 6} (ChamferDistancePytorch/chamfer{2D,5D,6D}:
 * chamfer6D.cu:11-183 forward, :185-237 backward; same contract as the 3-D entries above): xyz1 f32[B,N,D],
 * xyz2 f32[B,M,D] -> dist1 f32[B,N], dist2 f32[B,M] (squared), idx1 i32[B,N], idx2 i32[B,M]; ties -> lowest index.
 * Squared distance in the reference's contracted order (FMUL(d1,d1); FFMA(d0,d0,.); FFMA(d2,d2,.) ...).
 * backward accumulates into caller-zeroed gradxyz1 f32[B,N,D], gradxyz2 f32[B,M,D]. */
int pcfm_chamfer_nd_forward(const float* xyz1, const float* xyz2, int b, int n, int m, int d, float* dist1,
                            float* dist2, int* idx1, int* idx2, pcfm_stream_t stream);
int pcfm_chamfer_nd_backward(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                             const int* idx1, const int* idx2, int b, int n, int m, int d, float* gradxyz1,
                             float* gradxyz2, pcfm_stream_t stream);

/* cd[i,j] = mean_k dist1(S_i,R_j)[k] + mean_l dist2(S_i,R_j)[l]  with dist1/dist2 as pcfm_chamfer3d_forward.
 * S f32[ns,N,3], R f32[nr,M,3] -> cd f32[ns,nr].  Each distance is evaluated once and feeds both minima. */
int pcfm_chamfer3d_allpairs(const float* s, const float* r, int ns, int nr, int n, int m, float* cd,
                            pcfm_stream_t stream);
/* emd[i,j] = matchcost(approxmatch(S_i,R_j)) / N  (emd.py:48-50), never materialising `match`.
 * workspace: see pcfm_emd_allpairs_workspace. */
size_t pcfm_emd_allpairs_workspace(int ns, int nr, int n, int m);
int pcfm_emd_allpairs(const float* s, const float* r, int ns, int nr, int n, int m, float* emd,
                      void* workspace, size_t workspace_bytes, pcfm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PCFM_B200_H */
