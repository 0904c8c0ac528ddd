// cg.cuh -- internal declarations for the device-resident projected CG
#pragma once
#include "numeric.cuh"
