// oracle/cvshim: forwards to the minimal OpenCV stand-in (test infrastructure only).
#pragma once
#include "../../cvshim.h"
