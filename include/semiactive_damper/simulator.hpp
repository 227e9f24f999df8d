// drop-in name for semiactive_damper/simulator.hpp of the reference
#pragma once
#include "cgmres_b200/models.hpp"
typedef cgmres_b200::SemiactiveDamperSimulator Simulator;
