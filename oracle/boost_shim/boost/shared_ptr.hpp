// Minimal stand-in for <boost/shared_ptr.hpp>, used only to compile the reference's
// CPU path into oracle/_ref (test infrastructure; Boost is not installed in this image).
#pragma once
#include <memory>
namespace boost {
    using std::shared_ptr;
    using std::make_shared;
    using std::static_pointer_cast;
    using std::dynamic_pointer_cast;
}
