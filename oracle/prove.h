// ORACLE — TEST INFRASTRUCTURE ONLY (see fields.h header).
// (filled in by the prove-pipeline restatement: AIR evaluation, quotients, FRI, PoW, proof, verifier)
#pragma once
#include "vcs.h"
