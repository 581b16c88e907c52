// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// C-ABI shim around the reference's OWN CPU functions.  The function bodies are not
// stored in this repository: oracle/build_ref.sh slices them out of the read-only
// reference checkout at build time into oracle/_ref/*.inc (git-ignored):
//   nnsearch.inc    <- src/tf_ops/nn_distance/tf_nndistance.cpp:21-43
//   approxmatch.inc <- src/tf_ops/approxmatch/tf_approxmatch.cpp:23-140
// and this file is compiled against them into oracle/_ref/libtfops_ref_cpu.so.
#include <algorithm>
#include <vector>
#include <math.h>
#include <string.h>
#include "_ref/nnsearch.inc"
#include "_ref/approxmatch.inc"

extern "C" {
void ref_nnsearch(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx) {
    nnsearch(b, n, m, xyz1, xyz2, dist, idx);
}
void ref_approxmatch_cpu(int b, int n, int m, const float* xyz1, const float* xyz2, float* match) {
    approxmatch_cpu(b, n, m, xyz1, xyz2, match);
}
void ref_matchcost_cpu(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match, float* cost) {
    matchcost_cpu(b, n, m, xyz1, xyz2, match, cost);
}
// The reference zeroes only grad1's x component (tf_approxmatch.cpp:108-109, quirk Q4);
// pre-zero the whole buffer so y/z do not accumulate into uninitialised memory.
void ref_matchcostgrad_cpu(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match, float* grad1, float* grad2) {
    memset(grad1, 0, sizeof(float) * (size_t)b * n * 3);
    matchcostgrad_cpu(b, n, m, xyz1, xyz2, match, grad1, grad2);
}
}
