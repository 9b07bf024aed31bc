// host_preops.h — pure-host pieces of the engine pre-ops shared by the product library and the generator library.
#pragma once
#include <stdint.h>

#include "../../include/lorikeet_b200.h"

// initialize_pcr_error_model (...engine.rs:169-193): cache[n] for repeat lengths 0..100
void lk_engine_fill_pcr_cache(const lk_engine_args *args, uint8_t cache[101]);
