/* sepomp.h -- kept so that code including the reference's individual headers still builds;
 * everything lives in sep.h (seplib-b200). */
#include "sep.h"
