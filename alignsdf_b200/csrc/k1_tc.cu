// tcgen05 / TMEM decoder kernel -- placeholder entry points (filled in by the tensor-core milestone).
#include "common.cuh"

extern "C" int64_t asdf_tc_static_bytes(void) { return 0; }
extern "C" int64_t asdf_tc_sample_floats(void) { return 0; }
extern "C" int asdf_tc_eval(const asdf_tc_desc*, const void*, const float*, const asdf_query*, float*, float*,
                            int32_t*, void*) {
  asdf::set_error("asdf_tc_eval: tensor-core kernel not built in this revision");
  return ASDF_ERR_UNSUPPORTED;
}
