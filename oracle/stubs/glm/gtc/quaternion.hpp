// stand-in: see glm/glm.hpp in this directory
#pragma once
#include "glm/glm.hpp"
