#include "../../cvfull.hpp"
