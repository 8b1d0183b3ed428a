#include "oracle.hpp"
