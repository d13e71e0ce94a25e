// sdk_types.hpp -- the one SDK type that crosses the drop-in boundary.
//
// Inside the real driver tree the Slamtec SDK headers define
// sl_lidar_response_measurement_node_hq_t (reference src/sdk/include/sl_lidar_cmd.h:272-278);
// build with -DRPL_HAVE_SLAMTEC_SDK and this header just includes them.  Stand-alone (this
// repository, tests) the identical packed layout comes from the C-ABI header.
#pragma once
#include <cstdint>

#include "../../include/rpl_b200.h"

#ifdef RPL_HAVE_SLAMTEC_SDK
#include "sl_lidar.h"
#include "sl_lidar_driver.h"
static_assert(sizeof(sl_lidar_response_measurement_node_hq_t) == sizeof(rpl_node_hq), "node layout");
#else
using sl_lidar_response_measurement_node_hq_t = rpl_node_hq;
using sl_u8 = uint8_t;
using sl_u16 = uint16_t;
using sl_u32 = uint32_t;
using sl_result = uint32_t;
#endif
