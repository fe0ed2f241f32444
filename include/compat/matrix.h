// matrix.h of the reference (Matrix: val[i][j], operator<<) -> the slice the pipeline uses
#pragma once
#include "../goicp_dropin.hpp"
