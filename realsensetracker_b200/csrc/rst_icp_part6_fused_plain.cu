/* Slice 6 of the ICP kernel variants (fused_plain); see the end of rst_icp_kernels.inl. */
#define RST_ICP_PART 6
#include "rst_icp_kernels.inl"
