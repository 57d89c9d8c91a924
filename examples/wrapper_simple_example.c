/* A C caller of libgapb200.so (the role src/Programs/quip_wrapper_simple_example_C.c plays for libquip): a hydrogen pair in a periodic box,
 * first through the one-shot entry point gap_b200_wrapper_simple, then through the handle interface (initialise / cutoff / calc / finalise).
 * Build and run (needs a B200; there is no CPU fallback):
 *   gcc -Iinclude examples/wrapper_simple_example.c -Lquip_b200 -lgapb200 -Wl,-rpath,$PWD/quip_b200 -o /tmp/wrapper_simple_example
 *   /tmp/wrapper_simple_example tests/golden/GAP.xml
 */
#include <stdio.h>
#include <stdlib.h>

#include "gap_b200.h"

#define N_ATOMS 2

static int fail(const char* where) {
  fprintf(stderr, "%s: %s\n", where, gap_last_error());
  return 1;
}

int main(int argc, char** argv) {
  const char* xml = argc > 1 ? argv[1] : "gp.xml";
  const int species = argc > 2 ? atoi(argv[2]) : 1;
  /* Fortran layouts: cell(3,3) column-major (columns = cell vectors), positions(3,N) */
  const double cell[9] = {12.0, 0.0, 0.0, 0.0, 12.0, 0.0, 0.0, 0.0, 12.0};
  const double positions[3 * N_ATOMS] = {1.50, 2.25, 3.00, 2.35, 2.75, 3.40};
  const int numbers[N_ATOMS] = {species, species};
  const int n = N_ATOMS, periodic[3] = {1, 1, 1};
  double e_total = 0.0, forces[3 * N_ATOMS], stress_virial[9], e_atom[N_ATOMS], e_again = 0.0, forces_again[3 * N_ATOMS];

  if (gap_b200_wrapper_simple(xml, &n, cell, numbers, positions, &e_total, forces, stress_virial)) return fail("gap_b200_wrapper_simple");
  printf("Energy = %.12e\n", e_total);
  printf("Force0 = %.12e %.12e %.12e\n", forces[0], forces[1], forces[2]);

  gap_potential* pot = NULL;
  if (gap_potential_filename_initialise(&pot, "IP GAP", xml, 0)) return fail("gap_potential_filename_initialise");
  printf("Cutoff = %.6f\n", gap_potential_cutoff(pot));
  if (gap_potential_calc(pot, n, positions, numbers, cell, periodic, "", &e_again, e_atom, forces_again, NULL, NULL)) {
    gap_potential_finalise(pot);
    return fail("gap_potential_calc");
  }
  printf("Energy2 = %.12e\n", e_again);
  printf("LocalE = %.12e %.12e\n", e_atom[0], e_atom[1]);
  gap_potential_finalise(pot);
  return 0;
}
