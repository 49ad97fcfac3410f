/* Slice 0 of the ICP kernel variants (iter_plain); see the end of rst_icp_kernels.inl. */
#define RST_ICP_PART 0
#include "rst_icp_kernels.inl"
