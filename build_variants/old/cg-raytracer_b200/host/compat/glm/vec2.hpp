#pragma once
#include "vec3.hpp"
