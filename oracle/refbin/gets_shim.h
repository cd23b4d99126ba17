/* TEST INFRASTRUCTURE.  The reference's post-processing programs (postprocessing/plcm.c, pspt.c) are K&R C that reads its
 * answers with gets(), which glibc no longer declares.  This header is force-included (-include) when oracle/refbin/
 * refrun.py:build_plcm compiles the UNMODIFIED reference source into oracle/_ref/plcm; it only supplies gets(). */
#include <stdio.h>
#include <string.h>
static char *c2d_gets(char *s)
{
    if (!fgets(s, 1000, stdin)) { s[0] = 0; return 0; }
    s[strcspn(s, "\r\n")] = 0;
    return s;
}
#define gets c2d_gets
