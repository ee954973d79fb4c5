#include "cmpc_oracle.h"
