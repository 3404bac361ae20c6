// capi_internal.h -- the opaque handle of the inner C ABI
#pragma once
#include "engine.h"

struct q1t_state {
    q1t::DeviceVectorState *impl;
    bool owns;
};
