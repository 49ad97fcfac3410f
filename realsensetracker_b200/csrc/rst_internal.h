// rst_internal.h — what the optional engines (rst_icp3d.cu) need from the context of rst_capi.cu.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "rst_align.h"

namespace rst {
cudaStream_t ctx_stream(rst_ctx* c);
int ctx_device(rst_ctx* c);
void ctx_set_error(rst_ctx* c, const std::string& msg);
void ctx_count_launches(rst_ctx* c, int n);
// one extension slot per context: *slot is freed with free_fn(*slot) in rst_ctx_destroy
void** ctx_ext_slot(rst_ctx* c, void (***free_fn)(void*));
}  // namespace rst
