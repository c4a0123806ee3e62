// main.cpp -- process entry; exception -> exit code mapping of the reference's serial branch
// [REF src/main.cpp:46-65].
#include <exception>
#include <iostream>

void run_serial(int argc, char** argv);

int main(int argc, char** argv) {
  try {
    run_serial(argc, argv);
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
