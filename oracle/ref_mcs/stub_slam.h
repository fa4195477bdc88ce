// stub_slam.h -- data-only stand-ins for the three SLAM container classes that /root/reference/src/cORBmatcher.cpp reads
// (cMapPoint, cMultiFrame, cMultiKeyFrame), so that the reference's matcher compiles WHERE IT LIES without the rest of the
// system (g2o optimiser, DBoW2 vocabulary IO, OpenGV, Pangolin).   TEST INFRASTRUCTURE (oracle/Makefile target `ref`).
//
// This header is force-included (-include) ahead of the reference's own headers and pre-defines the include guards of
// include/cMapPoint.h, cMultiKeyFrame.h and cMultiFrame.h, so those three files are skipped; everything else the matcher uses
// is the reference's own code: cORBmatcher.{h,cpp}, cam_system_omni.{h,cpp} (WorldToCamHom_fast, MtMc bookkeeping),
// cam_model_omni.{h,cpp}, cConverter.{h,cpp}, misc.{h,cpp}, ThirdParty/DBoW2/DBoW2/FeatureVector.{h,cpp}.
//
// The classes below carry the public fields the matcher touches, under the reference's names and types (cited), filled by
// oracle/ref_mcs/wrap_match.cpp from flat arrays.  The only LOGIC in here is the grid lookup GetFeaturesInArea / PosInGrid --
// it belongs to cMultiFrame.cpp / cMultiKeyFrame.cpp, which cannot be compiled here, and is restated from
// src/cMultiFrame.cpp:272-353 and src/cMultiKeyFrame.cpp:694-737 -- and trivial accessors.  Map mutations the matcher performs
// (AddObservation / AddMapPoint / Replace) are recorded in call order so that tests can compare them.
#pragma once
#define MAPPOINT_H
#define MULTIKEYFRAME_H
#define MULTIFRAME_H

#include <opencv2/opencv.hpp>
#include <set>
#include <unordered_map>
#include <vector>

#include "DBoW2/DBoW2/FeatureVector.h"
#include "cam_system_omni.h"

#define FRAME_GRID_ROWS 48     // include/cMultiFrame.h:47
#define FRAME_GRID_COLS 64     // include/cMultiFrame.h:48

namespace MultiColSLAM
{
using std::vector;     // include/cORBmatcher.h:137 says `vector<...>` unqualified: the skipped headers leak a using-directive
class cMultiKeyFrame;
class cMapPoint;

// one recorded map mutation: kind 0 = pMP->AddObservation(pKF, idx) followed by pKF->AddMapPoint(pMP, idx), 1 = pMP->Replace(other)
struct RefMutation { int kind; int mp; int other_or_idx; };
struct RefMutationLog { std::vector<RefMutation> ops; };

class cMapPoint      // include/cMapPoint.h:44-161 (fields read by the matcher)
{
public:
	int id = -1;                                   // position in the caller's map-point array
	bool bad = false;
	cv::Vec3d worldPos, normal;
	double minDist = 0.0, maxDist = 0.0;
	int nObs = 0;
	std::vector<uint64_t> desc, dmask;
	std::unordered_map<cMultiKeyFrame*, std::vector<size_t>> obs;   // key frame -> keypoint indices
	RefMutationLog* log = nullptr;

	cv::Vec3d GetWorldPos() { return worldPos; }
	cv::Vec3d GetNormal() { return normal; }
	int TotalNrObservations() { return nObs; }
	bool isBad() { return bad; }
	const uint64_t* GetDescriptorPtr() { return desc.data(); }
	const uint64_t* GetDescriptorMaskPtr() { return dmask.data(); }
	double GetMinDistanceInvariance() { return minDist; }
	double GetMaxDistanceInvariance() { return maxDist; }
	bool IsInKeyFrame(cMultiKeyFrame* pKF) { return obs.count(pKF) != 0; }
	std::vector<size_t> GetIndexInKeyFrame(cMultiKeyFrame* pKF)
	{
		auto it = obs.find(pKF);
		return it != obs.end() ? it->second : std::vector<size_t>(1, (size_t)-1);
	}
	void AddObservation(cMultiKeyFrame* pKF, const size_t& idx)
	{
		obs[pKF].push_back(idx);
		if (log) log->ops.push_back(RefMutation{0, id, (int)idx});
	}
	void Replace(cMapPoint* pMP) { if (log) log->ops.push_back(RefMutation{1, id, pMP->id}); }

	// include/cMapPoint.h:101-105
	std::vector<double> mTrackProjX;
	std::vector<double> mTrackProjY;
	std::vector<bool> mbTrackInView;
	std::vector<int> mnTrackScaleLevel;
	std::vector<double> mTrackViewCos;
};

// the 64 x 48 bucket grid per camera (src/cMultiFrame.cpp:167-184, 342-353): contiguous keypoint ids in insertion order
struct RefGrid
{
	std::vector<std::vector<std::vector<std::vector<size_t>>>> g;      // [cam][col][row]
	std::vector<double> invW, invH;
	std::vector<int> minX, minY, maxX, maxY;
	void build(const std::vector<cv::KeyPoint>& keys, const std::unordered_map<size_t, int>& k2c, const std::vector<int>& w, const std::vector<int>& h)
	{
		const int nc = (int)w.size();
		g.assign(nc, std::vector<std::vector<std::vector<size_t>>>(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS)));
		invW.resize(nc); invH.resize(nc); minX.assign(nc, 0); minY.assign(nc, 0); maxX.resize(nc); maxY.resize(nc);
		for (int c = 0; c < nc; ++c)
		{
			maxX[c] = w[c]; maxY[c] = h[c];
			invW[c] = static_cast<double>(FRAME_GRID_COLS) / static_cast<double>(maxX[c] - minX[c]);     // src/cMultiFrame.cpp:136-137
			invH[c] = static_cast<double>(FRAME_GRID_ROWS) / static_cast<double>(maxY[c] - minY[c]);
		}
		for (size_t i = 0; i < keys.size(); ++i)
		{
			const int c = k2c.find(i)->second;
			const int px = cvRound((keys[i].pt.x - minX[c]) * invW[c]);                                   // PosInGrid :342-353
			const int py = cvRound((keys[i].pt.y - minY[c]) * invH[c]);
			if (px < 0 || px >= FRAME_GRID_COLS || py < 0 || py >= FRAME_GRID_ROWS) continue;
			g[c][px][py].push_back(i);
		}
	}
};

class cMultiFrame     // include/cMultiFrame.h:57-182
{
public:
	cMultiCamSys_ camSystem;
	std::vector<cv::KeyPoint> mvKeys;
	std::vector<cv::Vec3d> mvKeysRays;
	DBoW2::FeatureVector mFeatVec;
	std::vector<cv::Mat> mDescriptors;
	std::vector<cv::Mat> mDescriptorMasks;
	std::vector<cMapPoint*> mvpMapPoints;
	std::vector<bool> mvbOutlier;
	int mnScaleLevels = 0;
	std::vector<double> mvScaleFactors;
	std::unordered_map<size_t, int> keypoint_to_cam;
	std::unordered_map<size_t, int> cont_idx_to_local_cam_idx;
	RefGrid grid;

	cv::Matx<double, 4, 4> GetPose() { return camSystem.Get_M_t(); }

	// src/cMultiFrame.cpp:272-340
	std::vector<size_t> GetFeaturesInArea(const int& cam, const double& x, const double& y, const double& r,
		int minLevel = -1, int maxLevel = -1) const
	{
		std::vector<size_t> vIndices;
		const int nMinCellX = (int)floor((x - grid.minX[cam] - r) * grid.invW[cam]);
		const int minCellX = std::max(0, nMinCellX);
		if (minCellX >= FRAME_GRID_COLS) return vIndices;
		const int nMaxCellX = (int)ceil((x - grid.minX[cam] + r) * grid.invW[cam]);
		const int maxCellX = std::min(FRAME_GRID_COLS - 1, nMaxCellX);
		if (maxCellX < 0) return vIndices;
		const int nMinCellY = (int)floor((y - grid.minY[cam] - r) * grid.invH[cam]);
		const int minCellY = std::max(0, nMinCellY);
		if (minCellY >= FRAME_GRID_ROWS) return vIndices;
		const int nMaxCellY = (int)ceil((y - grid.minY[cam] + r) * grid.invH[cam]);
		const int maxCellY = std::min(FRAME_GRID_ROWS - 1, nMaxCellY);
		if (maxCellY < 0) return vIndices;
		bool bCheckLevels = true;
		bool bSameLevel = false;
		if (minLevel == -1 && maxLevel == -1) bCheckLevels = false;
		else if (minLevel == maxLevel) bSameLevel = true;
		for (int ix = minCellX; ix <= maxCellX; ++ix)
			for (int iy = minCellY; iy <= maxCellY; ++iy)
			{
				const std::vector<size_t>& vCell = grid.g[cam][ix][iy];
				for (size_t j = 0; j < vCell.size(); ++j)
				{
					const cv::KeyPoint& kp = mvKeys[vCell[j]];
					if (bCheckLevels && !bSameLevel) { if (kp.octave < minLevel || kp.octave > maxLevel) continue; }
					else if (bSameLevel) { if (kp.octave != minLevel) continue; }
					if (std::abs(kp.pt.x - x) > r || std::abs(kp.pt.y - y) > r) continue;
					vIndices.push_back(vCell[j]);
				}
			}
		return vIndices;
	}
};

class cMultiKeyFrame  // include/cMultiKeyFrame.h:49-313
{
public:
	cMultiCamSys_ camSystem;
	std::unordered_map<size_t, int> keypoint_to_cam;
	std::unordered_map<size_t, int> cont_idx_to_local_cam_idx;
	std::vector<cv::KeyPoint> mvKeys;
	std::vector<cv::Vec3d> mvKeysRays;
	std::vector<cv::Mat> mDescriptors, mDescriptorMasks;
	std::vector<cMapPoint*> mvpMapPoints;
	std::vector<double> mvScaleFactors;
	int mnScaleLevels = 0;
	DBoW2::FeatureVector mFeatVec;
	RefGrid grid;
	int id = -1;

	cv::Matx44d GetPose() { return camSystem.Get_M_t(); }                                       // src/cMultiKeyFrame.cpp:138-142
	cv::Matx44d GetPoseInverse() { return cConverter::invMat(camSystem.Get_M_t()); }             // :144-148
	cv::Vec3d GetCameraCenter() { const cv::Matx44d m = camSystem.Get_M_t(); return cv::Vec3d(m(0, 3), m(1, 3), m(2, 3)); }   // :150-156
	DBoW2::FeatureVector GetFeatureVector() { return mFeatVec; }
	void AddMapPoint(cMapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
	std::set<cMapPoint*> GetMapPoints()
	{
		std::set<cMapPoint*> s;
		for (size_t i = 0; i < mvpMapPoints.size(); ++i)
			if (mvpMapPoints[i] && !mvpMapPoints[i]->isBad()) s.insert(mvpMapPoints[i]);        // :273-285
		return s;
	}
	std::vector<cMapPoint*> GetMapPointMatches() { return mvpMapPoints; }
	cMapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
	cv::KeyPoint GetKeyPoint(const size_t& idx) const { return mvKeys[idx]; }
	cv::Vec3d GetKeyPointRay(const size_t& idx) const { return mvKeysRays[idx]; }
	int GetKeyPointScaleLevel(const size_t& idx) const { return mvKeys[idx].octave; }
	std::vector<cv::KeyPoint> GetKeyPoints() const { return mvKeys; }
	std::vector<cv::Vec3d> GetKeyPointsRays() const { return mvKeysRays; }
	const uint64_t* GetDescriptorRowPtr(const int& cam, const size_t& idx) const { return mDescriptors[cam].ptr<uint64_t>((int)idx); }
	const uint64_t* GetDescriptorMaskRowPtr(const int& cam, const size_t& idx) const { return mDescriptorMasks[cam].ptr<uint64_t>((int)idx); }
	std::vector<cv::Mat> GetAllDescriptors() const { return mDescriptors; }
	std::vector<cv::Mat> GetAllDescriptorMasks() const { return mDescriptorMasks; }
	double GetScaleFactor(int nLevel = 1) const { return mvScaleFactors[nLevel]; }
	std::vector<double> GetScaleFactors() const { return mvScaleFactors; }
	int GetScaleLevels() const { return mnScaleLevels; }

	// src/cMultiKeyFrame.cpp:694-737: no level filter, cell range clamped, `<=`-style window test as written there
	std::vector<size_t> GetFeaturesInArea(const int& cam, const double& x, const double& y, const double& r) const
	{
		std::vector<size_t> vIndices;
		const int nMinCellX = (int)floor((x - grid.minX[cam] - r) * grid.invW[cam]);
		const int minCellX = std::max(0, nMinCellX);
		if (minCellX >= FRAME_GRID_COLS) return vIndices;
		const int nMaxCellX = (int)ceil((x - grid.minX[cam] + r) * grid.invW[cam]);
		const int maxCellX = std::min(FRAME_GRID_COLS - 1, nMaxCellX);
		if (maxCellX < 0) return vIndices;
		const int nMinCellY = (int)floor((y - grid.minY[cam] - r) * grid.invH[cam]);
		const int minCellY = std::max(0, nMinCellY);
		if (minCellY >= FRAME_GRID_ROWS) return vIndices;
		const int nMaxCellY = (int)ceil((y - grid.minY[cam] + r) * grid.invH[cam]);
		const int maxCellY = std::min(FRAME_GRID_ROWS - 1, nMaxCellY);
		if (maxCellY < 0) return vIndices;
		for (int ix = minCellX; ix <= maxCellX; ++ix)
			for (int iy = minCellY; iy <= maxCellY; ++iy)
			{
				const std::vector<size_t>& vCell = grid.g[cam][ix][iy];
				for (size_t j = 0; j < vCell.size(); ++j)
				{
					const cv::KeyPoint& kp = mvKeys[vCell[j]];
					if (std::abs(kp.pt.x - x) <= r && std::abs(kp.pt.y - y) <= r) vIndices.push_back(vCell[j]);
				}
			}
		return vIndices;
	}
};
}
