#include "qt_publisher_shim.h"
