// multiple_controller/model1.hpp of the reference is mass_spring_damper/model.hpp renamed Model1
#pragma once
#include "cgmres_b200/models.hpp"
typedef cgmres_b200::MassSpringDamperModel Model1;
