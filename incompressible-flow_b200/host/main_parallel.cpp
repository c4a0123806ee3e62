// main_parallel.cpp -- process entry of the run_parallel-compatible driver; exception -> exit code mapping of the
// reference's MPI branch [REF src/main.cpp:17-44] (no MPI_Init here: placement comes from the environment).
#include <exception>
#include <iostream>

void run_parallel(int argc, char** argv);

int main(int argc, char** argv) {
  try {
    run_parallel(argc, argv);
  } catch (int i) {
    return i;
  } catch (const std::exception& e) {
    std::cout << "Exception\n" << e.what() << std::endl;
    return 1;
  } catch (...) {
    std::cout << "Unknown exception" << std::endl;
    return 1;
  }
  return 0;
}
