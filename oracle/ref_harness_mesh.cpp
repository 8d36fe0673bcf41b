// mesh / heightmap / broadphase entry points of the reference oracle (filled in below)
