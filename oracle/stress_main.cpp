// stress_main.cpp -- driver for the reference's OWN dynamic_emitter_stress_test() (tests/tests.cpp:434-514: 10 000 steps of adding,
// removing and replacing points in two variable-radius sets, every step compared with BruteforceNSearch), which the reference's
// main.cpp keeps behind `if (false)`.  Linked with the unmodified tests/tests.cpp + tests/BruteforceNSearch.cpp against the drop-in
// header include/TreeNSearch (oracle/Makefile: ref-tests).  Test infrastructure only.
#include "tests.h"

int main()
{
    dynamic_emitter_stress_test();
    return 0;
}
