/* The reference's C example (src/Programs/quip_wrapper_simple_example_C.c: two atoms in a 20 A box through quip_wrapper_simple_) against
 * libgapb200.so: the one-shot entry point gap_b200_wrapper_simple, then the handle interface (initialise / cutoff / calc / finalise) on the same
 * configuration.  Build and run (needs a B200; there is no CPU fallback):
 *   gcc -Iinclude examples/wrapper_simple_example.c -Lquip_b200 -lgapb200 -Wl,-rpath,$PWD/quip_b200 -o /tmp/wrapper_simple_example
 *   /tmp/wrapper_simple_example tests/golden/GAP.xml
 */
#include <stdio.h>
#include <stdlib.h>

#include "gap_b200.h"

int main(int argc, char** argv) {
  const char* xml = argc > 1 ? argv[1] : "gp.xml";
  int n = 2;
  double lattice[3][3] = {{20.0, 0.0, 0.0}, {0.0, 20.0, 0.0}, {0.0, 0.0, 20.0}};
  int Z[2] = {1, 1};
  double coord[2][3] = {{-7.110371, -3.533572, 2.147261}, {-7.933029, -3.234956, 2.573383}};
  double energy = 0.0, force[2][3], virial[3][3];
  if (argc > 2) Z[0] = Z[1] = atoi(argv[2]);

  if (gap_b200_wrapper_simple(xml, &n, &lattice[0][0], Z, &coord[0][0], &energy, &force[0][0], &virial[0][0])) {
    fprintf(stderr, "gap_b200_wrapper_simple: %s\n", gap_last_error());
    return 1;
  }
  printf("Energy = %.12e\n", energy);
  printf("Force0 = %.12e %.12e %.12e\n", force[0][0], force[0][1], force[0][2]);

  gap_potential* pot = NULL;
  if (gap_potential_filename_initialise(&pot, "IP GAP", xml, 0)) {
    fprintf(stderr, "gap_potential_filename_initialise: %s\n", gap_last_error());
    return 1;
  }
  int pbc[3] = {1, 1, 1};
  double e2 = 0.0, f2[2][3], local_e[2];
  printf("Cutoff = %.6f\n", gap_potential_cutoff(pot));
  if (gap_potential_calc(pot, n, &coord[0][0], Z, &lattice[0][0], pbc, "", &e2, local_e, &f2[0][0], NULL, NULL)) {
    fprintf(stderr, "gap_potential_calc: %s\n", gap_last_error());
    gap_potential_finalise(pot);
    return 1;
  }
  printf("Energy2 = %.12e\n", e2);
  printf("LocalE = %.12e %.12e\n", local_e[0], local_e[1]);
  gap_potential_finalise(pot);
  return 0;
}
