#pragma once
#include "pcl/point_types.h"
