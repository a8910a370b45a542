// see ../cvshim.hpp (type-only stand-in, test infrastructure)
#pragma once
#include <cvshim.hpp>
