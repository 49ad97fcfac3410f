/* Slice 5 of the ICP kernel variants (iter_photo_gm); see the end of rst_icp_kernels.inl. */
#define RST_ICP_PART 5
#include "rst_icp_kernels.inl"
