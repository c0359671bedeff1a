// stand-in header, see ../cvshim.hpp (oracle test infrastructure)
#pragma once
#include "cvshim.hpp"
