/* Slice 8 of the ICP kernel variants (fused_gm); see the end of rst_icp_kernels.inl. */
#define RST_ICP_PART 8
#include "rst_icp_kernels.inl"
