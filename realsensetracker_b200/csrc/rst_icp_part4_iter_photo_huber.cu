/* Slice 4 of the ICP kernel variants (iter_photo_huber); see the end of rst_icp_kernels.inl. */
#define RST_ICP_PART 4
#include "rst_icp_kernels.inl"
