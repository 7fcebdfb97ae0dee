// Stand-in for the TensorFlow header of the same path: see tests/tf_stub/tf_stub.h (test infrastructure).
#pragma once
#include "../../../tf_stub.h"
