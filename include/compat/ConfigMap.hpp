// ConfigMap.hpp of the reference -> the header-only ConfigMap of the drop-in
#pragma once
#include "../goicp_dropin.hpp"
