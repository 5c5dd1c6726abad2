// oracle/ stand-in for the CMake-generated header (asr_config_generated.h.in:18)
#pragma once
#define ASR_VERSION "0.2.0"
