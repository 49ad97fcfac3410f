/* Slice 3 of the ICP kernel variants (iter_photo); see the end of rst_icp_kernels.inl. */
#define RST_ICP_PART 3
#include "rst_icp_kernels.inl"
