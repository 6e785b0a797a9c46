// cta_solver.h -- CTA-per-trajectory path for large systems (declared here, defined in cta_solver.cu).
#pragma once
