#include "../cvfull.hpp"
