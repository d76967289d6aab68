// Circuit text -> DEM text, and a DEM shot sampler (see circuit.cc).
#pragma once
#include <cstdint>
#include <string>
#include <string_view>

#include "dem.h"

namespace tsr {

// Throws std::invalid_argument for instructions outside the supported subset.
std::string circuit_to_dem_text(std::string_view circuit_text);

// Samples `shots` i.i.d. shots from the error mechanisms of a (flattened) DEM into bit-packed rows
// (bit k of byte j = detector / observable 8j+k).  Buffers must be zeroed by the caller; `obs` may
// be null.  Deterministic in (dem, seed, shots).
void sample_dem_b8(const DemModel& flat, uint64_t seed, uint64_t shots, uint8_t* dets, uint64_t det_stride,
                   uint8_t* obs, uint64_t obs_stride);

}  // namespace tsr
