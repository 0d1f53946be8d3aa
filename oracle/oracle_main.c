/* TEST INFRASTRUCTURE — command-line front end of the C restatement oracle (same arguments as
 * oracle/ref_main.cpp so both outputs can be diffed).
 * usage: dftatom_oracle Z levels mixing rmax delta method [precision=6] [max_vcycles=100] */
#include <stdio.h>
#include <stdlib.h>
#include "dftatom_oracle.h"

int main(int argc, char** argv)
{
    if (argc < 7) { fprintf(stderr, "usage: %s Z levels mixing rmax delta method [precision] [max_vcycles]\n", argv[0]); return 2; }
    orc_options o;
    o.Z = atoi(argv[1]); o.levels = atoi(argv[2]); o.mixing = atof(argv[3]); o.max_r = atof(argv[4]);
    o.delta = atof(argv[5]); o.method = atoi(argv[6]);
    const int precision = argc > 7 ? atoi(argv[7]) : 6;
    const int max_vcycles = argc > 8 ? atoi(argv[8]) : 100;
    orc_scf_print(&o, precision, max_vcycles);
    return 0;
}
