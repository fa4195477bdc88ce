// reached as "../../config.h" from the reference's vendored g2o headers (cmake would generate it there): forwards to the
// config.h the reference ships under ThirdParty/g2o/g2o.  TEST INFRASTRUCTURE.
#include <g2o/config.h>
