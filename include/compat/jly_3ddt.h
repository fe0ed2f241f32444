// jly_3ddt.h of the reference (DT3D, CELL, EMPTYCELL) -> the B200 drop-in classes
#pragma once
#include "../goicp_dropin.hpp"
