#include "qt_zmq_shim.h"
