#pragma once
#include "../../include/starst3r_b200.h"
