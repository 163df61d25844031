{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE BangPatterns             #-}
-- | Thin FFI shim over libpwmscan.so (include/pwmscan.h).
--
-- NOT COMPILED IN THIS REPOSITORY'S ENVIRONMENT (no GHC in the build image): it is the binding a
-- maintainer of kaizhang/bioinformatics-toolkit adds so that Bio.Motif.Search / Bio.Motif /
-- Bio.Data.Bed.Utils keep their exported signatures while the work runs on a B200.  It is mirrored
-- 1:1 by the ctypes layer bioinformatics_toolkit_b200/_lib.py, which IS exercised by the tests.
module Bio.Motif.CUDA
    ( withDevice
    , findTFBSWithGPU
    , scoresGPU
    , maxMatchScGPU
    , pValueToScoreGPU
    , scoreCDFGPU
    , scanRegionsGPU
    , scanChunksGPU
    , hitsAffinityGPU
    , alignAllPairsGPU
    , PackedDNA (..)
    , packDNA
    , scanPackedChunksGPU
    , mkPlan
    , getSeqRaw
    , batchesOf
    ) where

import           Control.Exception        (bracket)
import           Control.Monad            (when)
import qualified Data.ByteString          as B
import qualified Data.ByteString.Unsafe   as BU
import           Data.Int
import qualified Data.Matrix.Unboxed      as M
import qualified Data.Vector.Storable     as S
import qualified Data.Vector.Storable.Mutable as SM
import qualified Data.Vector.Unboxed      as U
import           Data.Word
import           Foreign
import           Foreign.C.String
import           Foreign.C.Types
import           Data.IORef
import           System.Environment       (lookupEnv)
import           System.IO                (Handle, SeekMode (AbsoluteSeek), hSeek)
import           System.IO.Unsafe         (unsafePerformIO)

import           Bio.Motif                (Bkgd (..), CDF (..), PWM (..), size)
import           Bio.Seq                  (DNA, toBS)

data Ctx; data Seq; data Motifs; data Plan; data Hits; data AlnSet

-- `safe` so that a long scan does not block the GHC runtime.
foreign import ccall safe "pwmscan_ctx_create"       c_ctx_create   :: CInt -> Ptr (Ptr Ctx) -> IO CInt
foreign import ccall safe "pwmscan_ctx_destroy"      c_ctx_destroy  :: Ptr Ctx -> IO ()
foreign import ccall safe "pwmscan_last_error"       c_last_error   :: Ptr Ctx -> IO CString
foreign import ccall safe "pwmscan_motifs_create"    c_motifs_create :: Ptr Ctx -> Int32 -> Ptr Int32 -> Ptr Double -> Ptr Double -> Ptr (Ptr Motifs) -> IO CInt
foreign import ccall safe "pwmscan_motifs_set_sigma" c_set_sigma    :: Ptr Motifs -> Int32 -> CInt -> Ptr Double -> IO CInt
foreign import ccall safe "pwmscan_motifs_free"      c_motifs_free  :: Ptr Motifs -> IO ()
foreign import ccall safe "pwmscan_plan_create"      c_plan_create  :: Ptr Ctx -> Ptr Motifs -> Ptr Double -> CInt -> CInt -> Ptr (Ptr Plan) -> IO CInt
foreign import ccall safe "pwmscan_plan_free"        c_plan_free    :: Ptr Plan -> IO ()
foreign import ccall safe "pwmscan_scan"             c_scan         :: Ptr Ctx -> Ptr Seq -> Ptr Plan -> Ptr (Ptr Hits) -> IO CInt
foreign import ccall safe "pwmscan_scan_host"        c_scan_host    :: Ptr Ctx -> Ptr Plan -> Ptr Word8 -> Ptr Int64 -> Int32 -> CInt -> Ptr (Ptr Hits) -> IO CInt
foreign import ccall safe "pwmscan_hits_count"       c_hits_count   :: Ptr Hits -> IO Int64
foreign import ccall safe "pwmscan_hits_first_invalid" c_hits_first_invalid :: Ptr Hits -> CInt -> Ptr Int64 -> IO CInt
foreign import ccall safe "pwmscan_hits_columns"     c_hits_columns :: Ptr Hits -> Ptr (Ptr Int32) -> Ptr (Ptr Word8) -> Ptr (Ptr Int32) -> Ptr (Ptr Int64) -> Ptr (Ptr Double) -> IO CInt
foreign import ccall safe "pwmscan_hits_free"        c_hits_free    :: Ptr Hits -> IO ()
foreign import ccall safe "pwmscan_hits_affinity"    c_hits_affinity :: Ptr Ctx -> Ptr Hits -> Int32 -> Ptr Int64 -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Int64 -> IO CInt
foreign import ccall safe "pwmscan_scan_host_first_invalid" c_first_invalid :: Ptr Ctx -> CInt -> Ptr Int64 -> IO CInt
foreign import ccall safe "pwmscan_seq_upload"       c_seq_upload   :: Ptr Ctx -> Ptr Word8 -> Int64 -> CInt -> Ptr (Ptr Seq) -> IO CInt
foreign import ccall safe "pwmscan_seq_free"         c_seq_free     :: Ptr Seq -> IO ()
foreign import ccall safe "pwmscan_scores"           c_scores       :: Ptr Ctx -> Ptr Seq -> Ptr Motifs -> Int32 -> CInt -> Ptr Double -> Int64 -> IO CInt
foreign import ccall safe "pwmscan_max_match_sc"     c_max_match    :: Ptr Ctx -> Ptr Seq -> Ptr Motifs -> Ptr Double -> IO CInt
foreign import ccall safe "pwmscan_pvalue_to_score"  c_pvalue       :: Ptr Ctx -> Ptr Motifs -> Double -> Ptr Double -> IO CInt
foreign import ccall safe "pwmscan_score_cdf"        c_score_cdf    :: Ptr Ctx -> Ptr Motifs -> Int32 -> Ptr (Ptr Double) -> Ptr (Ptr Double) -> Ptr Int64 -> IO CInt
foreign import ccall safe "pwmscan_free"             c_free         :: Ptr a -> IO ()
foreign import ccall safe "pwmscan_alnset_create"    c_alnset_create :: Ptr Ctx -> Int32 -> Ptr Int32 -> Ptr Double -> CInt -> Double -> CInt -> Ptr (Ptr AlnSet) -> IO CInt
foreign import ccall safe "pwmscan_alnset_update"    c_alnset_update :: Ptr AlnSet -> Int32 -> Int32 -> Ptr Double -> IO CInt
foreign import ccall safe "pwmscan_alnset_pairs"     c_alnset_pairs  :: Ptr AlnSet -> Int64 -> Ptr Int32 -> Ptr Int32 -> Ptr Double -> Ptr Word8 -> Ptr Int32 -> IO CInt
foreign import ccall safe "pwmscan_alnset_free"      c_alnset_free   :: Ptr AlnSet -> IO ()
foreign import ccall safe "pwmscan_alnset_matrix_init" c_alnset_matrix_init :: Ptr AlnSet -> Ptr Int32 -> Ptr Int32 -> Ptr Double -> Ptr Word8 -> Ptr Int32 -> IO CInt
foreign import ccall safe "pwmscan_alnset_merge_step"  c_alnset_merge_step  :: Ptr AlnSet -> Int32 -> Int32 -> Int32 -> Ptr Double -> Ptr Int32 -> Ptr Int32 -> Ptr Double -> Ptr Word8 -> Ptr Int32 -> IO CInt
foreign import ccall safe "pwmscan_abi_version"       c_abi_version  :: IO CInt
foreign import ccall safe "pwmscan_last_timing"       c_last_timing  :: Ptr Ctx -> Ptr Double -> CInt -> IO CInt
foreign import ccall safe "pwmscan_seq_upload_segments" c_seq_upload_segments :: Ptr Ctx -> Ptr Word8 -> Ptr Int64 -> Int32 -> CInt -> Ptr (Ptr Seq) -> IO CInt
foreign import ccall safe "pwmscan_seq_length"        c_seq_length   :: Ptr Seq -> IO Int64
foreign import ccall safe "pwmscan_seq_first_invalid" c_seq_first_invalid :: Ptr Ctx -> Ptr Seq -> CInt -> Ptr Int64 -> IO CInt
foreign import ccall safe "pwmscan_motifs_get_sigma"  c_get_sigma    :: Ptr Motifs -> Int32 -> CInt -> Ptr Double -> IO CInt
foreign import ccall safe "pwmscan_motifs_count"      c_motifs_count :: Ptr Motifs -> IO Int32
foreign import ccall safe "pwmscan_plan_info"         c_plan_info    :: Ptr Plan -> Ptr Double -> CInt -> IO CInt
foreign import ccall safe "pwmscan_find_tfbs"         c_find_tfbs    :: Ptr Ctx -> Ptr Seq -> Ptr Motifs -> Ptr Double -> CInt -> CInt -> Ptr (Ptr Hits) -> IO CInt
foreign import ccall safe "pwmscan_cutoff_cdfs"       c_cutoff_cdfs  :: Ptr Ctx -> Ptr Motifs -> Double -> Ptr Double -> Ptr (Ptr Double) -> Ptr (Ptr Double) -> Ptr Int64 -> IO CInt
foreign import ccall safe "pwmscan_align_pairs"       c_align_pairs  :: Ptr Ctx -> Ptr Motifs -> Int64 -> Ptr Int32 -> Ptr Int32 -> CInt -> Double -> CInt -> Ptr Double -> Ptr Word8 -> Ptr Int32 -> IO CInt
foreign import ccall safe "pwmscan_format_bed6"       c_format_bed6  :: Int64 -> Ptr Int32 -> Ptr Int32 -> Ptr Word8 -> Ptr Int64 -> Ptr Int64 -> Ptr Int32 -> Ptr Int64
                                                                      -> CString -> Ptr Int64 -> CString -> Ptr Int64 -> Ptr Int32 -> Ptr CString -> Ptr Int64 -> IO CInt
foreign import ccall safe "pwmscan_format_dist"       c_format_dist  :: Int64 -> Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Double -> CString -> Ptr Int64 -> Ptr CString -> Ptr Int64 -> IO CInt
foreign import ccall safe "pwmscan_hits_compact"     c_hits_compact :: Ptr Hits -> Ptr (Ptr Word32) -> Ptr (Ptr Double) -> Ptr (Ptr Int64) -> Ptr (Ptr Int32) -> Ptr (Ptr Word16) -> IO CInt
-- hit lists of a multi-item scan arrive through a callback (item order, calling thread); the callee owns them
type HitsCb = Ptr () -> Int32 -> Ptr Hits -> IO ()
foreign import ccall "wrapper" mkHitsCb :: HitsCb -> IO (FunPtr HitsCb)
foreign import ccall safe "pwmscan_scan_many"        c_scan_many    :: Ptr Ctx -> Ptr Plan -> Int32 -> Ptr (Ptr Seq) -> FunPtr HitsCb -> Ptr () -> IO CInt
foreign import ccall safe "pwmscan_scan_host_many"   c_scan_host_many :: Ptr Ctx -> Ptr Plan -> Int32 -> Ptr (Ptr Word8) -> Ptr Int64 -> CInt -> FunPtr HitsCb -> Ptr () -> IO CInt
-- the packed genome store (2 bits + 1 mask bit per base, made once at mkIndex time; hostpack.cu)
foreign import ccall unsafe "pwmscan_packed_units"   c_packed_units :: Int64 -> Int64
foreign import ccall safe "pwmscan_pack_host"        c_pack_host    :: Ptr Word8 -> Int64 -> CInt -> Ptr Word32 -> Ptr Word32 -> Ptr Int64 -> Ptr Int64 -> CInt -> IO CInt
foreign import ccall safe "pwmscan_seq_upload_packed" c_seq_upload_packed :: Ptr Ctx -> Ptr Word32 -> Ptr Word32 -> Int64 -> Ptr (Ptr Seq) -> IO CInt
foreign import ccall safe "pwmscan_scan_packed_many" c_scan_packed_many :: Ptr Ctx -> Ptr Plan -> Int32 -> Ptr (Ptr Word32) -> Ptr (Ptr Word32) -> Ptr Int64 -> FunPtr HitsCb -> Ptr () -> IO CInt
foreign import ccall safe "pwmscan_align_all_pairs"  c_align_all    :: Ptr Ctx -> Ptr Motifs -> CInt -> Double -> CInt -> Ptr Double -> Ptr Word8 -> Ptr Int32 -> IO CInt

-- | The reference signals failure with `error` (pure code) — so do we.
check :: Ptr Ctx -> CInt -> IO ()
check ctx rc = when (rc /= 0) $ c_last_error ctx >>= peekCString >>= error

withDevice :: Int -> (Ptr Ctx -> IO a) -> IO a
withDevice dev = bracket open c_ctx_destroy
  where open = alloca $ \p -> do { rc <- c_ctx_create (fromIntegral dev) p; check nullPtr rc; peek p }

-- | One process-wide context (device 0, or PWMSCAN_DEVICE).  The pure entry points below go
-- through it with unsafePerformIO; every call is deterministic, so this is referentially transparent.
{-# NOINLINE theCtx #-}
theCtx :: Ptr Ctx
theCtx = unsafePerformIO $ do
    dev <- maybe 0 read <$> lookupEnv "PWMSCAN_DEVICE"
    alloca $ \p -> do { rc <- c_ctx_create (fromIntegral (dev :: Int)) p; check nullPtr rc; peek p }

-- U.Vector / M.Matrix are unpinned: copy to Storable before the call.
withMotifs :: Bkgd -> [PWM] -> (Ptr Motifs -> IO a) -> IO a
withMotifs (BG (a, c, g, t)) pwms act =
    S.unsafeWith ncol $ \pn -> S.unsafeWith flat $ \pm -> S.unsafeWith bg $ \pb ->
    bracket (alloca $ \pp -> do { rc <- c_motifs_create theCtx (fromIntegral $ length pwms) pn pm pb pp; check theCtx rc; peek pp })
            c_motifs_free act
  where
    ncol = S.fromList $ map (fromIntegral . size) pwms :: S.Vector Int32
    flat = S.concat $ map (S.convert . M.flatten . _mat) pwms :: S.Vector Double
    bg   = S.fromList [a, c, g, t]

-- | Body of findTFBS / findTFBSWith (Bio/Motif/Search.hs:25-58): hits (position, score), ascending.
findTFBSWithGPU :: Maybe (U.Vector Double) -> Bkgd -> PWM -> DNA a -> Double -> Bool -> [(Int, Double)]
findTFBSWithGPU sigma bg pwm dna thres skip = unsafePerformIO $
    withMotifs bg [pwm] $ \mo -> do
        case sigma of
            Just sg -> S.unsafeWith (S.convert sg) $ \ps -> c_set_sigma mo 0 0 ps >>= check theCtx
            Nothing -> return ()
        bracket (with thres $ \pt -> alloca $ \pp -> do
                    rc <- c_plan_create theCtx mo pt 1 (if skip then 1 else 0) pp
                    check theCtx rc >> peek pp) c_plan_free $ \plan ->
          BU.unsafeUseAsCStringLen (toBS dna) $ \(p, n) ->
          withArray [0, fromIntegral n] $ \poff ->
          bracket (alloca $ \ph -> do
                      rc <- c_scan_host theCtx plan (castPtr p) poff 1 0 ph
                      check theCtx rc >> peek ph) c_hits_free $ \h -> do
              cnt <- fromIntegral <$> c_hits_count h
              alloca $ \ppos -> alloca $ \psc -> do
                  _ <- c_hits_columns h nullPtr nullPtr nullPtr ppos psc
                  pos <- peek ppos >>= peekArray cnt
                  sc  <- peek psc  >>= peekArray cnt
                  return $! zip (map fromIntegral pos) sc

-- | scores (Bio/Motif.hs:127-145).
scoresGPU :: Bkgd -> PWM -> DNA a -> [Double]
scoresGPU bg pwm dna = unsafePerformIO $ withMotifs bg [pwm] $ \mo ->
    BU.unsafeUseAsCStringLen (toBS dna) $ \(p, n) ->
    bracket (alloca $ \ps -> do { rc <- c_seq_upload theCtx (castPtr p) (fromIntegral n) 2 ps; check theCtx rc; peek ps }) c_seq_free $ \sq -> do
        let nwin = max 0 (n - size pwm + 1)
        allocaArray (max 1 nwin) $ \out -> do
            c_scores theCtx sq mo 0 0 out (fromIntegral nwin) >>= check theCtx
            peekArray nwin out

-- | maxMatchSc (Bio/Motif/Search.hs:98-109).
maxMatchScGPU :: Bkgd -> PWM -> DNA a -> Double
maxMatchScGPU bg pwm dna = unsafePerformIO $ withMotifs bg [pwm] $ \mo ->
    BU.unsafeUseAsCStringLen (toBS dna) $ \(p, n) ->
    bracket (alloca $ \ps -> do { rc <- c_seq_upload theCtx (castPtr p) (fromIntegral n) 2 ps; check theCtx rc; peek ps }) c_seq_free $ \sq ->
    alloca $ \out -> c_max_match theCtx sq mo out >>= check theCtx >> peek out

-- | pValueToScore (Bio/Motif.hs:213-214), batched.
pValueToScoreGPU :: Double -> Bkgd -> [PWM] -> [Double]
pValueToScoreGPU p bg pwms = unsafePerformIO $ withMotifs bg pwms $ \mo ->
    allocaArray (length pwms) $ \out -> c_pvalue theCtx mo p out >>= check theCtx >> peekArray (length pwms) out

-- | scoreCDF (Bio/Motif.hs:282-326).
scoreCDFGPU :: Bkgd -> PWM -> CDF
scoreCDFGPU bg pwm = unsafePerformIO $ withMotifs bg [pwm] $ \mo ->
    alloca $ \px -> alloca $ \pp -> alloca $ \pn -> do
        c_score_cdf theCtx mo 0 px pp pn >>= check theCtx
        n <- fromIntegral <$> peek pn
        x <- peek px; q <- peek pp
        xs <- peekArray n x; ps <- peekArray n q
        c_free x >> c_free q
        return $ CDF $ U.fromList $ zip xs ps

-- | Per-batch body of scanMotif (Bio/Data/Bed/Utils.hs:101-118): regions laid end to end, one
-- segment each; returns (region index, motif index, isForward, position in region, score) in the
-- reference's order (region, motif, forward hits ascending, reverse hits ascending).
scanRegionsGPU :: Ptr Plan -> B.ByteString -> [Int] -> IO [(Int, Int, Bool, Int, Double)]
scanRegionsGPU plan bytes lens =
    BU.unsafeUseAsCStringLen bytes $ \(p, _) ->
    withArray (map fromIntegral $ scanl (+) 0 lens) $ \poff ->
    bracket (alloca $ \ph -> do
                rc <- c_scan_host theCtx plan (castPtr p) poff (fromIntegral $ length lens) 1 ph   -- 1 = upper-case like fromBS
                check theCtx rc >> peek ph) c_hits_free $ \h -> do
        cnt <- fromIntegral <$> c_hits_count h
        alloca $ \pm -> alloca $ \pst -> alloca $ \psg -> alloca $ \ppos -> alloca $ \psc -> do
            _ <- c_hits_columns h pm pst psg ppos psc
            ms <- peek pm >>= peekArray cnt; st <- peek pst >>= peekArray cnt; sg <- peek psg >>= peekArray cnt
            ps <- peek ppos >>= peekArray cnt; sc <- peek psc >>= peekArray cnt
            return [ (fromIntegral g, fromIntegral m, s == 0, fromIntegral q, v) | (g, m, s, q, v) <- zip5' sg ms st ps sc ]
  where zip5' (a:as) (b:bs) (c:cs) (d:ds) (e:es) = (a, b, c, d, e) : zip5' as bs cs ds es
        zip5' _ _ _ _ _ = []

-- | Whole-chromosome / whole-genome scans: the (motif block x chromosome chunk) items of the reference's findTFBS'
-- (Bio/Motif/Search.hs:60-84), several of them in flight on the device (pwmscan_scan_host_many).  Every chunk is a
-- strict ByteString of the genome store (`getChrom`, Bio/Seq/IO.hs:77-84, cut into pieces that overlap by
-- (longest motif - 1) bytes); the result of chunk i is, per (motif, strand), the (position in chunk, score) pairs
-- in ascending position — what `findTFBS` yields for that PWM on that chunk.
scanChunksGPU :: Ptr Plan -> Int -> [B.ByteString] -> IO [[((Int, Bool), [(Int, Double)])]]
scanChunksGPU plan nMotif chunks = do
    acc <- newIORef []
    cb  <- mkHitsCb $ \_ _ h -> hitsRuns nMotif h >>= \r -> c_hits_free h >> modifyIORef' acc (r :)
    let go [] ptrs lens = withArray (reverse ptrs) $ \pa -> withArray (reverse lens) $ \pl ->
            c_scan_host_many theCtx plan (fromIntegral $ length chunks) pa pl 1 cb nullPtr >>= check theCtx
        go (c : cs) ptrs lens = BU.unsafeUseAsCStringLen c $ \(p, n) -> go cs (castPtr p : ptrs) (fromIntegral n : lens)
    go chunks [] []
    freeHaskellFunPtr cb
    reverse <$> readIORef acc

-- | The plan `scanMotif` scans with (Bio/Data/Bed/Utils.hs:101-118): every CutoffMotif's PWM on both strands, ambiguous
-- bases skipped, its stored cutoff.  `parts` = [(pwm, cutoff)] in the order of the motif list (the reverse-complement
-- tables and both sigma vectors are computed by the library and are bit-identical to `_sigma` / `_sigmaRc`; pass a
-- CutoffMotif as (_motifPwm m, _cutoff m)).  Lives as long as the process (scanMotif builds it once per motif list).
mkPlan :: Bkgd -> [(PWM, Double)] -> IO (Ptr Plan)
mkPlan (BG (a, c, g, t)) parts =
    S.unsafeWith ncol $ \pn -> S.unsafeWith flat $ \pm -> S.unsafeWith bg $ \pb -> withArray (map snd parts) $ \pt -> do
        mo <- alloca $ \pp -> c_motifs_create theCtx (fromIntegral $ length parts) pn pm pb pp >>= check theCtx >> peek pp
        alloca $ \pp -> c_plan_create theCtx mo pt 2 1 pp >>= check theCtx >> peek pp
  where
    ncol = S.fromList $ map (fromIntegral . size . fst) parts :: S.Vector Int32
    flat = S.concat $ map (S.convert . M.flatten . _mat . fst) parts :: S.Vector Double
    bg   = S.fromList [a, c, g, t]

-- | `getSeq` (Bio/Seq/IO.hs:65-75) without `fromBS`: the bytes of (chr, start, end) as they are in the genome store —
-- the device upper-cases them and reports fromBS's verdict for the batch (pwmscan_scan_host_first_invalid).
-- Arguments are the three fields of the reference's `Genome` record (handle, index table lookup, header size).
getSeqRaw :: Handle -> (B.ByteString -> Maybe (Int, Int)) -> Int -> (B.ByteString, Int, Int) -> IO (Either String B.ByteString)
getSeqRaw h index headerSize (chr, start, end) = case index chr of
    Just (chrStart, chrSize)
        | end > chrSize -> return $ Left $ "Bio.Seq.getSeq: out of index: " ++ show end ++ ">" ++ show chrSize
        | otherwise     -> do hSeek h AbsoluteSeek $ fromIntegral $ headerSize + chrStart + start
                              Right <$> B.hGet h (end - start)
    _ -> return $ Left $ "Bio.Seq.getSeq: Cannot find " ++ show chr

-- | Lists of up to n consecutive elements (the batches of BED records scanMotif hands to the device); plain list version
-- of what the conduit in INTEGRATION.md does with `Data.Conduit.List.chunksOf`-style buffering.
batchesOf :: Int -> [a] -> [[a]]
batchesOf _ [] = []
batchesOf n xs = let (a, b) = splitAt n xs in a : batchesOf n b

-- | The hits of a single-sequence scan as runs: per (motif, strand) the (position, score) pairs in ascending position.
hitsRuns :: Int -> Ptr Hits -> IO [((Int, Bool), [(Int, Double)])]
hitsRuns nMotif h = do
    cnt <- fromIntegral <$> c_hits_count h
    alloca $ \pp -> alloca $ \ps -> alloca $ \pr -> do
        _ <- c_hits_compact h pp ps pr nullPtr nullPtr
        if cnt == 0 then return [] else do
            pos <- peek pp >>= peekArray cnt; sc <- peek ps >>= peekArray cnt
            ro  <- peek pr >>= peekArray (2 * nMotif + 1)
            let runs = zip ro (tail ro)
            return [ ((c `div` 2, even c), [ (fromIntegral q, v) | (q, v) <- take (fromIntegral (b - a)) $ drop (fromIntegral a) $ zip pos sc ])
                   | (c, (a, b)) <- zip [0 ..] runs, b > a ]

-- | The packed planes of a sequence (what `mkIndex` can store next to the bytes): (2-bit words, mask words, length).
-- Made once on the host; `packDNA` errors like `fromBS` (Bio/Seq.hs:86-95) on a byte outside the IUPAC alphabet.
data PackedDNA = PackedDNA !(S.Vector Word32) !(S.Vector Word32) !Int

packDNA :: B.ByteString -> PackedDNA
packDNA bs = unsafePerformIO $ BU.unsafeUseAsCStringLen bs $ \(p, n) -> do
    let units = fromIntegral (c_packed_units (fromIntegral n))
    s2 <- SM.new (2 * units); mk <- SM.new units
    bad <- SM.unsafeWith s2 $ \p2 -> SM.unsafeWith mk $ \pm -> alloca $ \pa -> alloca $ \pi' -> do
        _ <- c_pack_host (castPtr p) (fromIntegral n) 1 p2 pm pa pi' 0
        peek pi'
    when (bad >= 0) $ error "Bio.Seq.fromBS: unknown character"
    PackedDNA <$> S.unsafeFreeze s2 <*> S.unsafeFreeze mk <*> pure n

-- | `scanChunksGPU` for chunks of packed sequences: (planes, start, end) with start a multiple of 32 — 0.375 bytes per
-- base cross PCIe instead of 1.  Same results.
scanPackedChunksGPU :: Ptr Plan -> Int -> [(PackedDNA, Int, Int)] -> IO [[((Int, Bool), [(Int, Double)])]]
scanPackedChunksGPU plan nMotif chunks = do
    acc <- newIORef []
    cb  <- mkHitsCb $ \_ _ h -> hitsRuns nMotif h >>= \r -> c_hits_free h >> modifyIORef' acc (r :)
    let go [] p2s pms lens = withArray (reverse p2s) $ \pa -> withArray (reverse pms) $ \pb -> withArray (reverse lens) $ \pl ->
            c_scan_packed_many theCtx plan (fromIntegral $ length chunks) pa pb pl cb nullPtr >>= check theCtx
        go ((PackedDNA s2 mk _, st, en) : cs) p2s pms lens = S.unsafeWith s2 $ \p2 -> S.unsafeWith mk $ \pm ->
            go cs (p2 `plusPtr` (4 * (st `div` 16)) : p2s) (pm `plusPtr` (4 * (st `div` 32)) : pms) (fromIntegral (en - st) : lens)
    go chunks [] [] []
    freeHaskellFunPtr cb
    reverse <$> readIORef acc

-- | scanMotif's per-hit BED score on the device (Bio/Data/Bed/Utils.hs:109-123): `toAffinity (1 - cdf c sc)` for
-- every hit of `h`, given every motif's truncated CDF (the `_cdf` fields of the CutoffMotifs) as one
-- concatenated (offsets, scores, probabilities) triple.  Same integers as the Haskell expression.
hitsAffinityGPU :: Ptr Hits -> [CDF] -> IO [Int]
hitsAffinityGPU h cdfs = do
    cnt <- fromIntegral <$> c_hits_count h
    let pts  = [ U.toList v | CDF v <- cdfs ]
        offs = scanl (+) 0 (map length pts)
    withArray (map fromIntegral offs) $ \poff ->
      withArray (concatMap (map fst) pts) $ \px -> withArray (concatMap (map snd) pts) $ \py ->
      allocaArray (max 1 cnt) $ \pa -> do
        c_hits_affinity theCtx h (fromIntegral $ length cdfs) poff px py nullPtr pa >>= check theCtx
        map fromIntegral <$> peekArray cnt (pa :: Ptr Int64)

-- | All i<j alignments for `alignmentBy jsd (pFn gap) combFn` (Bio/Motif/Alignment.hs:83-132),
-- packed upper triangle in the order iterativeMerge initialises its matrix (Bio/Motif/Merge.hs:161-165).
alignAllPairsGPU :: Int -> Double -> Int -> [PWM] -> [(Double, (Bool, Int))]
alignAllPairsGPU penalKind gap combKind pwms = unsafePerformIO $ withMotifs (BG (0.25, 0.25, 0.25, 0.25)) pwms $ \mo ->
    let n = length pwms; np = n * (n - 1) `div` 2 in
    allocaArray (max 1 np) $ \pd -> allocaArray (max 1 np) $ \ps -> allocaArray (max 1 np) $ \pp -> do
        c_align_all theCtx mo (fromIntegral penalKind) gap (fromIntegral combKind) pd ps pp >>= check theCtx
        d <- peekArray np pd; s <- peekArray np ps; q <- peekArray np pp
        return [ (x, (y /= 0, fromIntegral z)) | (x, y, z) <- zip3 d s (q :: [Int32]) ]
