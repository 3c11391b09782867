// Empty stand-in: reference src/polyassembly.cpp:8 includes <jellyfish/mer_dna.hpp> but uses no
// Jellyfish symbol.  TEST INFRASTRUCTURE (oracle build only).
#pragma once
