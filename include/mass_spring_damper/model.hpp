// drop-in name for mass_spring_damper/model.hpp of the reference
#pragma once
#include "cgmres_b200/models.hpp"
typedef cgmres_b200::MassSpringDamperModel Model;
