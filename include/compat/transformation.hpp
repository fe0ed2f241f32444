// transformation.hpp of the reference (Transformation, point, point4D, the `properties` colour codes)
#pragma once
#include "../goicp_dropin.hpp"
using namespace std;   // as the reference header does (transformation.hpp:19)
enum properties {OG = 8204959, N = 30894, O = 15219528, NZ = 15231913, CZ = 4646984, CA = 16741671, DU = 7566712, OD1 = 0, C = 1 };   // transformation.hpp:36
