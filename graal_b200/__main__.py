"""python -m graal_b200 <data set folder> ...  (graal_b200.simulation.main)"""
import sys

from .simulation import main

sys.exit(main())
