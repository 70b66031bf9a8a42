// boost::filesystem stand-in — TEST INFRASTRUCTURE ONLY.
// The reference (matching/main.cpp:21, matching/matcher.cpp:21) uses boost::filesystem for
// directory iteration, path stem/extension and create_directory; Boost is not in this image.
// std::filesystem has the same semantics for every member the reference calls, including the
// quoted operator<< used when the score CSVs are written (matcher.cpp:203, :320).
#ifndef LAFIS_ORACLE_BOOST_FS_STANDIN
#define LAFIS_ORACLE_BOOST_FS_STANDIN
#include <filesystem>
namespace boost {
namespace filesystem = std::filesystem;
}
#endif
