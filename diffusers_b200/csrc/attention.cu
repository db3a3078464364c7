// placeholder until the tcgen05 attention kernel lands
#include "host_common.h"
namespace b200 { int init_attention() { return 0; } }
