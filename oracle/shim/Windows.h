// oracle/shim/Windows.h -- TEST INFRASTRUCTURE ONLY. ransmt.h:4 spells the header with a capital W.
#include "windows.h"
