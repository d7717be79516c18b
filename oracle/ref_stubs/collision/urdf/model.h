// ORACLE (test infrastructure only).  See urdf_model/model.h.
#pragma once
#include <urdf_model/model.h>
