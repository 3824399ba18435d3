/* Force-included (-include) when compiling the reference's Applications/TRAcT/tube.c on Linux.
 * glibc's <signal.h> declares gsignal(); syn_structs.h:29 declares `double gsignal;`.  Rename the
 * reference's symbol after pulling in the system header.  <sys/param.h> provides MAXPATHLEN. */
#include <signal.h>
#include <sys/param.h>
#define gsignal trm_ref_gsignal
