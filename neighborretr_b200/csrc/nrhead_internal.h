// Internal glue: pulls in the public C ABI so kernels and entry points share its constants.
#pragma once
#include "../../include/nrhead.h"
