// Shared bits of the two clustering command lines (directory scan in readdir order, name filter).
#ifndef SCEMA_B200_CLI_COMMON_H
#define SCEMA_B200_CLI_COMMON_H
#include <dirent.h>
#include <string>
#include <sys/types.h>
#include <vector>

namespace cli {

// Every directory entry in the order the filesystem hands them out ("." and ".." included, as the
// reference's read_directory does: clustering/mpi_comparison_test.cc:28-36). The batch order, and
// with it the partner order inside each result file, follows this enumeration.
inline bool list_directory(const std::string &dir, std::vector<std::string> &entries)
{
    DIR *d = opendir(dir.c_str());
    if (!d) return false;
    while (struct dirent *e = readdir(d)) entries.push_back(e->d_name);
    closedir(d);
    return true;
}

inline bool is_strain_file(const std::string &name) { return name.compare(0, 7, "strain_") == 0; }

// atoi of the name with its first "strain_" removed (mpi_comparison_test.cc:85-87)
inline unsigned id_from_name(std::string name)
{
    size_t pos = name.find("strain_");
    if (pos != std::string::npos) name.erase(pos, 7);
    return (unsigned)atoi(name.c_str());
}

}  // namespace cli
#endif
