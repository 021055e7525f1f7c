/* TEST INFRASTRUCTURE ONLY (oracle/).  The reference's command-line driver with the one change INTEGRATION.md
 * section 1 describes: reference src/tol.cpp cannot be compiled here (CPython 2), so this file plays its part
 * -- same positional arguments (src/arguments.cpp:32-46), same sequence as mission_select (src/tol.cpp:5-36:
 * build the problem object, runSNOPT(), writeJSON("snopt_results.json")) -- around the UNMODIFIED reference
 * objects problem*.o, snoptProblem.o, parameters.o, arguments.o, jsoncpp.o.
 *
 *   tol_dropin_ref   links the reference's own DefineFG.o           -> DEFINEGusrfg_ = reference CPU path
 *   tol_dropin_cuda  links libtolcuda instead (-DTOL_DROPIN_CUDA)    -> DEFINEGusrfg_ = sm_100a kernels,
 *                    plus the three binding lines (create_from_files, bind_global, destroy)
 * SNOPT is oracle/snmock.cpp in both.  Usage:  tol_dropin_* E N U Eg Ng Ug Rg aircraft mission [root/] */
#include <cstdio>
#include <iostream>

#include "problemG7.h"
#include "problemS10.h"
#ifdef TOL_DROPIN_CUDA
#include "tolcuda.h"
#endif

problem *prob = NULL; /* reference src/tol.cpp:3 */

int main(int argc, char *argv[]) {
    if (argc < 10) {
        std::cerr << "usage: " << argv[0] << " E N U Eg Ng Ug Rg aircraft mission [root/]" << std::endl;
        return 2;
    }
    arguments args(argv);
    if (argc > 10) args.root_path = argv[10];
    if (args.mission == "G7") prob = new problemG7(args);
    else if (args.mission == "S10") prob = new problemS10(args);
    else return 2;
#ifdef TOL_DROPIN_CUDA
    tolcuda_handle dev = NULL;
    int rc = tolcuda_create_from_files(args.root_path.c_str(), args.aircraft.c_str(), args.mission.c_str(), args.east,
                                       args.north, args.up, args.east_goal, args.north_goal, args.up_goal,
                                       args.radius_goal, /*ts_override=*/0, /*device=*/0, &dev);
    if (rc) {
        std::cerr << "tolcuda: " << tolcuda_last_error() << std::endl;
        return 1;
    }
    tolcuda_bind_global(dev);
#endif
    prob->runSNOPT();
    prob->writeJSON("snopt_results.json");
#ifdef TOL_DROPIN_CUDA
    tolcuda_destroy(dev);
#endif
    delete prob;
    return 0;
}
