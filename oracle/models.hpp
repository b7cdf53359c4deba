// The BASELINE config models live in tools/bench_models.hpp (shared by the fixture generator, the reference-arm
// timing tool, the integration test and the plugin benchmark tools/cuda_bench.cpp); kept here as a forwarding header.
#pragma once
#include "../tools/bench_models.hpp"
