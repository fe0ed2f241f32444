// jly_goicp.h of the reference (GoICP, POINT3D, ROTNODE, TRANSNODE) -> the B200 drop-in classes
#pragma once
#include "../goicp_dropin.hpp"
