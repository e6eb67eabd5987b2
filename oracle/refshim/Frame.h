// oracle/refshim -- TEST INFRASTRUCTURE ONLY: stands in for the reference header of the same name (see refshim.h).
#pragma once
#include "refshim.h"
