#include "thrust_emu.h"
