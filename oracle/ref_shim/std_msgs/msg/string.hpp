#pragma once
#include "ref_shim_msgs.hpp"
