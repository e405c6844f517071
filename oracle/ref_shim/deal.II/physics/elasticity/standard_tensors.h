#include <deal.II/base/dealii_min.h>
