// ORACLE (test infrastructure only).  See kdl/frames.hpp.
#pragma once
#include <kdl/frames.hpp>
