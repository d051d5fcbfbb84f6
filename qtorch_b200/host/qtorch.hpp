// qtorch_b200/host/qtorch.hpp -- umbrella header (counterpart of /root/reference/src/qtorch.hpp:21-32).
// Unlike the reference's headers these are include-guarded AND multi-TU safe (everything is inline).
#pragma once
#include "Timer.h"
#include "Exceptions.h"
#include "DeviceEngine.h"
#include "Node.h"
#include "Wire.h"
#include "Network.h"
#include "LineGraph.h"
#include "ContractionTools.h"
#include "leviParser.hpp"
#include "preprocess.h"
#include "Slicing.h"
#include "PlanCache.h"
using namespace qtorch;
