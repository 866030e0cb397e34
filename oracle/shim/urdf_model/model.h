// TEST INFRASTRUCTURE: stand-in for <urdf_model/model.h>
#pragma once
#include "../urdf/model.h"
