/*
 * include/monopsr_b200_tfops.h -- C ABI of the B200-native point-set ops.
 *
 * Drop-in boundary for the reference's custom-op kernel launchers
 * (kujason/monopsr, src/tf_ops).  Each entry point below replaces one C++-linkage
 * launcher that the reference's TF op shells bind to:
 *
 *   mpb_nn_distance        <- NmDistanceKernelLauncher      (tf_nndistance.cpp:168, tf_nndistance_g.cu:128-131)
 *   mpb_nn_distance_grad   <- NmDistanceGradKernelLauncher  (tf_nndistance.cpp:208, tf_nndistance_g.cu:152-157)
 *   mpb_approxmatch        <- approxmatchLauncher           (tf_approxmatch.cpp:141, tf_approxmatch_g.cu:180-182)
 *   mpb_matchcost          <- matchcostLauncher             (tf_approxmatch.cpp:142, tf_approxmatch_g.cu:226-228)
 *   mpb_matchcostgrad      <- matchcostgradLauncher         (tf_approxmatch.cpp:143, tf_approxmatch_g.cu:292-295)
 *
 * Contract (same as the reference launchers unless stated):
 *   - all pointers are DEVICE pointers to dense row-major fp32 / int32 buffers that the
 *     CALLER allocated (outputs and `temp` scratch included); the library owns nothing
 *     the caller can see.  Internal stream-ordered workspaces (cudaMallocAsync) are
 *     taken and released inside a call.
 *   - argument lists are the reference's, with `void* stream` (a cudaStream_t; NULL =
 *     legacy default stream, which is what the reference launches on) appended, and
 *     an int status returned instead of void: 0 = ok, >0 = the cudaError_t of the
 *     failing runtime call / launch, -1 = invalid argument (b,n,m <= 0 with non-empty
 *     output, NULL pointer).  The reference performs shape validation in the op shell
 *     (OP_REQUIRES, tf_nndistance.cpp:51-58,94-105; tf_approxmatch.cpp:150-158,219);
 *     the host-side Python mirror does the same checks.
 *   - b == 0 or an empty cloud side is accepted and is a no-op (status 0).
 *   - asynchronous, stateless, re-entrant across streams.
 *
 * Layouts: xyz (b,pts,3) AoS; dist (b,pts) f32; idx (b,pts) i32;
 *          match (b,m,n) with element [l,k] at l*n+k (query-major, the GPU layout,
 *          tf_approxmatch.py:21; quirk Q3 in SURVEY.md section 8).
 */
#ifndef MONOPSR_B200_TFOPS_H_
#define MONOPSR_B200_TFOPS_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Library identification: returns a static string "monopsr_b200 <ver> sm_100a". */
const char* mpb_version(void);

/* Bidirectional nearest neighbour (squared L2, first index wins ties).
 * result/result_i : (b,n) for xyz -> xyz2 ; result2/result2_i : (b,m) for xyz2 -> xyz. */
int mpb_nn_distance(int b, int n, const float* xyz, int m, const float* xyz2,
                    float* result, int* result_i, float* result2, int* result2_i,
                    void* stream);

/* Gradient of sum(dist1*grad_dist1)+sum(dist2*grad_dist2) w.r.t. both clouds.
 * Zeroes grad_xyz1/grad_xyz2 itself, like the reference launcher (tf_nndistance_g.cu:153-154). */
int mpb_nn_distance_grad(int b, int n, const float* xyz1, int m, const float* xyz2,
                         const float* grad_dist1, const int* idx1,
                         const float* grad_dist2, const int* idx2,
                         float* grad_xyz1, float* grad_xyz2, void* stream);

/* Approximate-EMD soft assignment.  temp: caller scratch of b*(n+m)*2 floats
 * (tf_approxmatch.cpp:167-170); may be NULL -- the B200 kernels keep that state on chip. */
int mpb_approxmatch(int b, int n, int m, const float* xyz1, const float* xyz2,
                    float* match, float* temp, void* stream);

/* cost[b] = sum_{l,k} ||xyz2[l]-xyz1[k]|| * match[l,k] */
int mpb_matchcost(int b, int n, int m, const float* xyz1, const float* xyz2,
                  const float* match, float* out, void* stream);

/* d cost / d xyz1 -> grad1 (b,n,3), d cost / d xyz2 -> grad2 (b,m,3); match is constant. */
int mpb_matchcostgrad(int b, int n, int m, const float* xyz1, const float* xyz2,
                      const float* match, float* grad1, float* grad2, void* stream);

/* Number of kernel launches issued by this library since load (all entry points);
 * used by bench.py to report "gpu_launches". */
unsigned long long mpb_launch_count(void);

/* CRC-32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start): the per-tensor checksum of TensorFlow
 * checkpoints (monopsr_b200/core/tf_checkpoint.py). */
unsigned mpb_crc32c(const void* data, unsigned long long n, unsigned crc);

#ifdef __cplusplus
}
#endif
#endif /* MONOPSR_B200_TFOPS_H_ */
