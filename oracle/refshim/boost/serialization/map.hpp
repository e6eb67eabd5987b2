// oracle/refshim -- TEST INFRASTRUCTURE ONLY: empty stand-in (see serialization.hpp).
#pragma once
#include "serialization.hpp"
