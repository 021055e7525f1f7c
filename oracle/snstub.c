/* TEST INFRASTRUCTURE ONLY (oracle/): empty bodies for the Fortran SNOPT entry points declared
 * in /root/reference/include/snopt/snopt.h:93-216.  libsnopt7 is commercial and absent; these
 * exist only so that the unmodified reference src/snoptProblem.cpp links.  The oracle never
 * calls runSNOPT(), so none of them is ever executed. */
#define SNSTUB(name) void name(void) {}
SNSTUB(f_sninit) SNSTUB(f_snspec) SNSTUB(f_sngetc) SNSTUB(f_sngeti) SNSTUB(f_sngetr)
SNSTUB(f_snset) SNSTUB(f_snseti) SNSTUB(f_snsetr) SNSTUB(f_snsetprint) SNSTUB(f_snend)
SNSTUB(f_snopta) SNSTUB(f_snkera) SNSTUB(f_snjac) SNSTUB(f_snmema) SNSTUB(f_snoptb)
SNSTUB(f_snkerb) SNSTUB(f_snoptc) SNSTUB(f_snkerc) SNSTUB(f_snmem)
