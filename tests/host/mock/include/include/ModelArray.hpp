/* forwarding header: the nextsimdg tree includes its headers as "include/X.hpp" */
#include "../ModelArray.hpp"
