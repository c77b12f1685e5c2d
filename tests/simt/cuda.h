// TEST INFRASTRUCTURE: what `#include <cuda.h>` resolves to when the kernels are built for the SIMT emulator.
#pragma once
#include "simt.h"
