// TEMPORARY stubs (replaced by net.cu).
#include "common.cuh"
#define EXPORT extern "C" __attribute__((visibility("default")))
EXPORT int prg_net_create(prg_net**, int, const void*, size_t, int, int, int) { prg::set_error("not implemented"); return PRG_ERR_STATE; }
EXPORT void prg_net_destroy(prg_net*) {}
EXPORT size_t prg_net_device_bytes(const prg_net*) { return 0; }
EXPORT int prg_unet_forward(prg_net*, const float*, const int64_t*, const float*, float*, int, prg_stream_t) { prg::set_error("not implemented"); return PRG_ERR_STATE; }
EXPORT int prg_maskunet_forward(prg_net*, const float*, float*, uint8_t*, float, int, prg_stream_t) { prg::set_error("not implemented"); return PRG_ERR_STATE; }
EXPORT int prg_sampler_run(prg_net*, const prg_sched*, int, const int*, int, float, const float*, const float*, const float*, uint64_t, int, float*, int, prg_stream_t) { prg::set_error("not implemented"); return PRG_ERR_STATE; }
