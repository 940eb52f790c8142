/* see asoundlib.h in this directory: the stub declares the external-plugin API there */
#include "asoundlib.h"
