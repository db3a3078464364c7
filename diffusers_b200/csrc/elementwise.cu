#include "host_common.h"
