// Host-side configuration logic (see astc_host_config.cpp).
#pragma once
#include "../../include/astcenc.h"
#include "astc_dev_tables.h"
#include "astc_host_tables.h"

namespace astc_host {
astcenc_error validate_cpu_float();
astcenc_error validate_config(astcenc_config& config);
astcenc_error config_init(astcenc_profile profile, unsigned int block_x, unsigned int block_y, unsigned int block_z, float quality, unsigned int flags,
                          astcenc_config* config);
// The device-side view of a validated config, including the dB -> squared error conversion.
void make_device_config(const astcenc_config& config, DevConfig& out);
}
