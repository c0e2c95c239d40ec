#!/bin/bash
# md5 of the SASS (addresses and encodings stripped) of every object given: two builds whose kernels are the same code
# print the same sums.  Used to check that a refactor left the GPU-measured kernels untouched.
for f in "$@"; do
  echo "$(cuobjdump -sass $f | grep -E '^\s+/\*[0-9a-f]{4}\*/|Function' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/\/\*.*//' | md5sum | cut -c1-16)  $f"
done
