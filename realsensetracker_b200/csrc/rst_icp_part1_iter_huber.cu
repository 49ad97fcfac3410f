/* Slice 1 of the ICP kernel variants (iter_huber); see the end of rst_icp_kernels.inl. */
#define RST_ICP_PART 1
#include "rst_icp_kernels.inl"
