// TEST INFRASTRUCTURE: forwards to the OpenCV stand-in used to build oracle/_ref (see cvstub.h).
#include "../../cvstub.h"
