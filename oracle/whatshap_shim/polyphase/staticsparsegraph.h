// TEST INFRASTRUCTURE.  Only the NodeId typedef of WhatsHap's StaticSparseGraph is used by the
// reference (src/alignmentstoreadset.cpp:607,674).
#pragma once
#include <cstdint>
class StaticSparseGraph { public: typedef uint32_t NodeId; };
