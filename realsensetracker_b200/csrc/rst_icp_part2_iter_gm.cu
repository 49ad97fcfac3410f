/* Slice 2 of the ICP kernel variants (iter_gm); see the end of rst_icp_kernels.inl. */
#define RST_ICP_PART 2
#include "rst_icp_kernels.inl"
