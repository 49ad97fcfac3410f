/* Slice 7 of the ICP kernel variants (fused_huber); see the end of rst_icp_kernels.inl. */
#define RST_ICP_PART 7
#include "rst_icp_kernels.inl"
