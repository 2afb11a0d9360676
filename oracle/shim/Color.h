/* shim: color.cpp:1 includes "Color.h" but the file is color.h */
#include "color.h"
