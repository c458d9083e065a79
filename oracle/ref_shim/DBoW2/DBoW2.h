#pragma once
#include "BowVector.h"
#include "FeatureVector.h"
