#include "qt_decode_shim.h"
