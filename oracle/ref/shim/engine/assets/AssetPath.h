// Build shim for the engine's asset-path helper (only Exists()/string() are used by
// reference Dataset.cpp:194-199).  Paths are taken as given.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <filesystem>
#include <string>

class AssetPathAbs
{
public:
	explicit AssetPathAbs(const std::string& p) : m_Path(p) {}
	bool Exists() const { return std::filesystem::exists(m_Path); }
	std::string string() const { return m_Path.string(); }
private:
	std::filesystem::path m_Path;
};
