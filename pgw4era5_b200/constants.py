"""Physical constants of the PGW path; values identical to the reference's
``constants.py:3-7`` (COSMO data_constants) because they enter the results bit
for bit."""
CON_RD = 287.05      # gas constant of dry air [J kg-1 K-1]
CON_G = 9.80665      # gravitational acceleration [m s-2]
CON_MW_MD = 0.622    # molar-mass ratio water vapour / dry air [1]
