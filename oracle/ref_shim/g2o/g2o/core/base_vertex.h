// stand-in for Thirdparty/g2o/g2o/core/base_vertex.h: sft_types.h includes it by this path; everything is in g2o_shim.h
#include "../../../inc/g2o_shim.h"
