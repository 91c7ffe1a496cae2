// placeholder until the wavefront integrator lands (next milestone)
#include "context.h"
namespace spb {
void renderStateDestroy(spb_ctx*) {}
}  // namespace spb
